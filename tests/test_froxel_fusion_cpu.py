"""The fused froxel column path (plainrenderer_b200/csrc/froxel_inc.cuh: the phases froxelColumnKernel runs per thread) on the CPU against
the oracle's four separate passes (froxelVolumeMaterial -> froxelLightScattering -> volumeLightingReprojection ->
volumetricLightingIntegration, Volumetrics.cpp:151-246), bit for bit.

The phase functions are host+device; tests/emul/froxel_fusion_host.cu loops over the thread indices of a block per phase, so what runs
here is the statement sequence of the kernel: the per-column view directions, the per-block depth tables, the binary16 rounding between
the stages and the order of the running sums. The GPU suite (test_zz_single_pass_gpu.py::test_froxel_passes_bit_exact, fused=True, and
every frame test with pass fusion) holds the kernel itself against the oracle."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import passes
from conftest import ROOT
from test_froxels_numpy import scene


@pytest.fixture(scope="module")
def host_lib():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not found: the host build of froxel_inc.cuh needs it")
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libfroxel_fusion_host.so"
    src = ROOT / "tests" / "emul" / "froxel_fusion_host.cu"
    csrc = ROOT / "plainrenderer_b200" / "csrc"
    deps = [src] + list(csrc.glob("*.cuh")) + list(csrc.glob("*.h")) + list((ROOT / "include").glob("*.h"))
    if not lib.exists() or lib.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        # host code under the contract's flags: no contraction, hardware fma for the explicit fmaf
        subprocess.run([nvcc, "-shared", "-std=c++17", "-O2", "-fmad=false", "--extended-lambda", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-fno-fast-math", "-diag-suppress", "177,550", "-I%s" % (ROOT / "include"), "-I%s" % csrc,
                        "-o", str(lib), str(src)], check=True, capture_output=True)
    h = C.CDLL(str(lib))
    h.froxel_fused_host.restype = C.c_int
    h.froxel_fused_host.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 4
    return h


def fused_on_host(h, ffi, res, noise, shadow, light_matrix2, light, settings, history, g, rows=None, z_lanes=64):
    w, hh, d = res
    lm = np.zeros((4, 16), np.float32)
    lm[2] = np.asarray(light_matrix2, np.float32)
    info = passes._shadow_cascade_info(ffi, lm)
    noise = np.ascontiguousarray(noise, np.uint8)
    shadow = np.ascontiguousarray(shadow, np.uint16)
    light = np.ascontiguousarray(light, np.float32)
    settings = np.ascontiguousarray(settings, np.float32)
    hist = np.ascontiguousarray(history, np.float16)
    gbytes = np.frombuffer(bytes(g), np.uint8).copy()
    outs = [np.zeros((d, hh, w, 4), np.float16) for _ in range(4)]
    y0, y1 = rows or (0, hh)
    rc = h.froxel_fused_host(w, hh, d, noise.ctypes.data, noise.shape[0], shadow.ctypes.data, shadow.shape[0], info.ctypes.data, light.ctypes.data, settings.ctypes.data,
                             hist.ctypes.data, gbytes.ctypes.data, y0, y1, z_lanes, *[o.ctypes.data for o in outs])
    assert rc == 0
    return outs


@pytest.mark.parametrize("z_lanes", [64, 32, 16])
@pytest.mark.parametrize("res,moving,cut,noise_size", [((12, 7, 16), False, False, 8), ((10, 6, 8), True, False, 8), ((9, 5, 8), True, True, 8), ((19, 9, 64), True, False, 8),
                                                       ((8, 3, 70), True, False, 8), ((11, 6, 16), True, False, 6), ((9, 4, 24), False, False, 32)])
def test_fused_froxel_columns_equal_the_oracles_four_passes(ffi, oracle, host_lib, res, moving, cut, noise_size, z_lanes):
    """noise_size 6: the generic repeat addressing of the density noise; 8 / 32: the power-of-two masks (froxelNoiseSample)"""
    cam, prev, noise, shadow, L, settings, light, history, sun = scene(res[0] * 10 + res[2], res, moving)
    if noise_size != noise.shape[0]:
        noise = np.random.default_rng(noise_size).integers(0, 256, (noise_size,) * 3, dtype=np.uint8)
    history[1, 2, 3, :] = np.nan
    history[2, 1, 0, 3] = np.inf
    want = passes.froxels(ffi, oracle, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    got = fused_on_host(host_lib, ffi, res, noise, shadow, L.T.ravel(), light, settings, history, want[4], z_lanes=z_lanes)
    for name, x, y in zip(("material", "scattering", "reprojection", "integration"), got, want):
        assert np.array_equal(x.view(np.uint16), y.view(np.uint16)), "%s: %d of %d halves differ" % (name, int((x.view(np.uint16) != y.view(np.uint16)).sum()), x.size)


def test_fused_froxel_columns_row_window(ffi, oracle, host_lib):
    """row sharding: a launch over rows [y0, y1) writes exactly those rows, with the whole-volume values"""
    res = (11, 8, 16)
    cam, prev, noise, shadow, L, settings, light, history, sun = scene(5, res, True)
    want = passes.froxels(ffi, oracle, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun)
    got = fused_on_host(host_lib, ffi, res, noise, shadow, L.T.ravel(), light, settings, history, want[4], rows=(2, 5))
    for x, y in zip(got, want):
        assert np.array_equal(x.view(np.uint16)[:, 2:5], y.view(np.uint16)[:, 2:5])
        assert not x.view(np.uint16)[:, :2].any() and not x.view(np.uint16)[:, 5:].any()
