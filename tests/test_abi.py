"""CPU-side checks of the drop-in boundary: the product library loads, exports every symbol the headers declare, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared(header, macro):
    text = (ROOT / "include" / header).read_text()
    return sorted(set(re.findall(r"%s\((\w+)\)\(" % macro, text)))


def test_headers_declare_the_reference_interface():
    fns = declared("plain_b200.h", "PLAIN_FN")
    # the compute-pass subset of RenderBackend's public section (RenderBackend.h:33-110)
    for name in ["create_image", "create_temporary_image", "create_uniform_buffer", "create_storage_buffer", "create_sampler", "create_compute_pass",
                 "update_compute_pass_shader_description", "set_global_descriptor_set_resources", "new_frame", "set_compute_pass_execution",
                 "prepare_for_drawcall_recording", "set_uniform_buffer_data", "set_storage_buffer_data", "render_frame", "resize_images", "get_image_description",
                 "get_image_global_texture_array_index", "get_swapchain_input_image", "get_renderpass_timings"]:
        assert name in fns


def test_product_library_exports_every_declared_symbol(product_lib):
    lib = C.CDLL(str(product_lib))
    for name in declared("plain_b200.h", "PLAIN_FN"):
        assert hasattr(lib, "plain_" + name), "libplain_b200.so does not export plain_" + name
    for name in declared("plain_frontend.h", "PLAIN_FE"):
        assert hasattr(lib, "plain_frontend_" + name), "libplain_b200.so does not export plain_frontend_" + name


def test_fast_contract_library_exports_the_same_abi(product_lib):
    """libplain_b200_fast.so: same C-ABI; its floating-point passes really are the SFU build (MUFU.EX2 / LG2 in
    the shading object) while the integer / LUT / rasterisation objects are shared with the exact library."""
    import plainrenderer_b200 as pr
    assert pr.LIB_FAST_PATH.exists()
    lib = C.CDLL(str(pr.LIB_FAST_PATH))
    for name in declared("plain_b200.h", "PLAIN_FN"):
        assert hasattr(lib, "plain_" + name)
    for name in declared("plain_frontend.h", "PLAIN_FE"):
        assert hasattr(lib, "plain_frontend_" + name)
    syms = subprocess.run(["nm", "-D", "--defined-only", str(pr.LIB_FAST_PATH)], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


def test_product_library_does_not_contain_the_oracle(product_lib):
    syms = subprocess.run(["nm", "-D", "--defined-only", str(product_lib)], capture_output=True, text=True).stdout
    assert "oracle_" not in syms
    needed = subprocess.run(["readelf", "-d", str(product_lib)], capture_output=True, text=True).stdout
    assert "liboracle" not in needed


def test_ffi_symbol_lists_match_headers(ffi):
    assert sorted(ffi.BACKEND_SYMBOLS) == declared("plain_b200.h", "PLAIN_FN")
    assert sorted(ffi.FRONTEND_SYMBOLS) == declared("plain_frontend.h", "PLAIN_FE")


def test_struct_sizes(ffi):
    assert C.sizeof(ffi.GlobalShaderInfo) == 340  # global.inc:4-33 std140
    assert C.sizeof(ffi.ImageDesc) == 36
    assert C.sizeof(ffi.ComputePassExecution) == 4 + 4 + 5 * 16 + 8 + 4 + 12 + 4 or C.sizeof(ffi.ComputePassExecution) % 8 == 0


def test_no_cpu_fallback_without_a_device(cuda):
    """Without a CUDA device backend_create must fail (non-zero) instead of falling back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    p = C.c_void_p()
    rc = cuda.b["backend_create"](C.c_int(0), C.c_uint32(64), C.c_uint32(64), C.byref(p))
    assert rc != 0 and not p.value


def test_unknown_shader_is_an_error(oracle, ffi):
    be = ffi.Backend(oracle, width=8, height=8)
    with pytest.raises(ffi.ApiError):
        be.create_compute_pass("noSuchShader.comp")
    be.close()
