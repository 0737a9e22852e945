"""The oracle's passes against the reference's OWN compute shaders, main() included.

oracle/_ref/liboracle_refmain.so is the oracle with the shaders listed in oracle/ref/ref_shaders.txt executed by the reference's GLSL: the
.comp files and the include files they name are compiled as C++ from where they lie under /root/reference (oracle/ref/glsl_shader_to_cpp.py
adapts spelling only and turns the `layout(...)` interface declarations into variables; arithmetic and texel access are oracle/glsl.h and
oracle/image.h, the contract both sides share) and registered as overrides of the oracle's restatements. The same frames rendered through
liboracle.so and through that library must agree in every image and buffer, bit for bit: that pins the statement order of every listed
main(), where tests/test_oracle_vs_reference_glsl.py pins the include functions one by one.

/root/reference does not travel: where neither the library nor the reference exists the tests skip (CPU suite only; no GPU test needs it)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import passes
from conftest import ROOT, Sequence, assert_snapshots_equal

LIB = ROOT / "oracle" / "_ref" / "liboracle_refmain.so"
_COMPLETED = set()  # the frame tests that ran in THIS process (the run counters of the library are per process: xdist may split the module)


@pytest.fixture(scope="module")
def refmain(ffi, oracle):
    if Path("/root/reference/resources/shaders").exists():
        subprocess.run(["bash", str(ROOT / "oracle" / "build_ref.sh")], check=True, capture_output=True)  # up to date: returns at once
    if not LIB.exists():
        pytest.skip("oracle/_ref/liboracle_refmain.so not built and /root/reference not present")
    api = ffi.Api(str(LIB), "oracle_", "oracle_frontend_")
    h = C.CDLL(str(LIB))
    h.oracle_refmain_shaders.restype = C.c_char_p
    api.refmain_runs = h.oracle_refmain_runs
    api.refmain_runs_of = lambda shader: h.oracle_refmain_runs_of(shader.encode())
    api.refmain_shaders = h.oracle_refmain_shaders().decode().split()
    return api


def listed():
    return [l.strip() for l in (ROOT / "oracle" / "ref" / "ref_shaders.txt").read_text().splitlines() if l.strip() and not l.startswith("#")]


def test_every_listed_shader_is_compiled_in(refmain):
    assert refmain.refmain_shaders == listed()
    assert len(refmain.refmain_shaders) >= 10


@pytest.mark.parametrize("res,moving,cut", [((12, 7, 16), False, False), ((10, 6, 8), True, False), ((9, 5, 8), True, True), ((19, 9, 64), True, False)])
def test_froxel_shaders(ffi, oracle, refmain, res, moving, cut):
    from test_froxels_numpy import scene
    cam, prev, noise, shadow, L, settings, light, history, sun = scene(res[0] * 10 + res[2], res, moving)
    history[1, 2, 3, :] = np.nan
    before = refmain.refmain_runs()
    a = passes.froxels(ffi, oracle, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    b = passes.froxels(ffi, refmain, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    assert refmain.refmain_runs() - before == 4, "the four froxel passes did not go through the reference's main()"
    for name, x, y in zip(("material", "scattering", "reprojection", "integration"), a, b):
        assert np.array_equal(x.view(np.uint16), y.view(np.uint16)), name


@pytest.mark.parametrize("w,h,frames,moving,instances,settings", [
    (120, 72, 4, True, 14, {}),
    (100, 60, 3, True, 10, dict(taa_use_clipping=0, taa_use_motion_vector_dilation=0, taa_use_separate_supersampling=1)),  # + temporalSupersampling.comp, colorToLuminance.comp
    (96, 56, 2, False, 10, dict(sdf_debug_mode=2)),  # sdfDebugVisualisation.comp
])
def test_frames_through_the_reference_shaders(ffi, oracle, refmain, w, h, frames, moving, instances, settings):
    """whole frame sequences (every history fed back): the oracle against the oracle with the listed passes run by the reference's GLSL"""
    try:
        a, b = Sequence(ffi, oracle, w, h, instances, **settings), Sequence(ffi, refmain, w, h, instances, **settings)
    except Exception as e:  # a setting this frontend does not know
        pytest.skip(str(e))
    before = refmain.refmain_runs()
    try:
        for f in range(frames):
            inputs = a.step(moving=moving)
            b.step(moving=moving, inputs=inputs)
            assert_snapshots_equal(b.snapshot(), a.snapshot(), "frame %d of %dx%d through the reference's shaders" % (f, w, h))
    finally:
        a.close()
        b.close()
    assert refmain.refmain_runs() - before >= 10 * frames, "the frames did not go through the reference's main()s"
    _COMPLETED.add(("frames", w, h))


@pytest.mark.parametrize("settings", [
    dict(diffuse_brdf=0, direct_multiscatter=1, taa_history_sampling_tech=1, half_res_trace=0),
    dict(diffuse_brdf=1, direct_multiscatter=2, use_geometry_aa=0, taa_history_sampling_tech=2, strict_influence_radius_cutoff=0),
    dict(diffuse_brdf=3, direct_multiscatter=3, taa_history_sampling_tech=3, sun_shadow_cascade_count=4, strict_influence_radius_cutoff=1),
    dict(indirect_lighting_tech=1, taa_filter_use_tonemapping=0, taa_history_sampling_tech=4, sun_direction_deg=(200.0, 80.0)),
])
def test_setting_variants_through_the_reference_shaders(ffi, oracle, refmain, settings):
    """the specialisation-constant variants of the shaders: the four diffuse BRDFs and multiscatter lobes of triangle.frag (+ geometric AA off,
    constant ambient), the bicubic history samplers of the TAA resolve, full-resolution trace, strict influence cut-off, 4 shadow cascades"""
    a, b = Sequence(ffi, oracle, 88, 52, 9, **settings), Sequence(ffi, refmain, 88, 52, 9, **settings)
    try:
        for f in range(3):
            inputs = a.step(moving=True)
            b.step(moving=True, inputs=inputs)
            assert_snapshots_equal(b.snapshot(), a.snapshot(), "frame %d with %s through the reference's shaders" % (f, settings))
    finally:
        a.close()
        b.close()
    _COMPLETED.add(("variants", tuple(sorted(settings))))


def test_frames_from_meshes_through_the_reference_shaders(ffi, oracle, refmain):
    """frames rendered end to end from `.plain` meshes (raster_inputs = 1): the depth prepass resolves its fragments through the reference's
    depthPrepass.frag main() (motion vectors, encoded normal) and every later pass through its own shader - moving camera, so the motion
    vectors are not zero"""
    from conftest import PlainSceneSequence
    from plainrenderer_b200 import assets
    a = PlainSceneSequence(ffi, oracle, assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_"), 112, 64)
    b = PlainSceneSequence(ffi, refmain, assets.Assets(LIB, "oracle_asset_"), 112, 64)
    try:
        for f in range(3):
            a.step(moving=True)
            b.step(moving=True)
            sa = a.snapshot()
            assert_snapshots_equal(b.snapshot(), sa, "frame %d from meshes through the reference's shaders" % f)
        assert sa["motion%d/0" % ((a.frame - 1) % 3)].any() or sa["motion0/0"].any() or sa["motion1/0"].any(), "the motion vectors of a moving camera are all zero"
    finally:
        a.close()
        b.close()
    _COMPLETED.add(("meshes",))


def test_zz_every_listed_shader_ran(refmain):
    """after the tests above: each listed shader was executed by the reference's main() at least once (none fell back to the oracle's restatement)"""
    if len(_COMPLETED) < 3 + 4 + 1:
        pytest.skip("needs the frame tests of this module in the same process (%d of 8 ran here)" % len(_COMPLETED))
    idle = [s for s in refmain.refmain_shaders + ["triangle.frag", "depthPrepass.frag"] if refmain.refmain_runs_of(s) == 0]  # the fragment shaders: behind oracle/shading_hook.h
    assert not idle, "never executed through the reference's main(): %s" % idle


def test_block_layouts_of_the_converter_match_the_c_abi_structs(ffi):
    """the std140 / std430 offsets oracle/ref/glsl_shader_to_cpp.py computes for the reference's interface blocks against the C structs of
    include/plain_frame_types.h (the layouts the backend's buffers actually have): the global uniform block member by member, the sizes of
    the others"""
    import re
    import sys
    shaders = Path("/root/reference/resources/shaders")
    if not shaders.exists():
        pytest.skip("/root/reference not present")
    sys.path.insert(0, str(ROOT / "oracle" / "ref"))
    import glsl_shader_to_cpp as conv
    text = conv.gather(shaders, "sdfDiffuseTrace.comp", set()) + conv.gather(shaders, "skyLut.comp", set()) + conv.gather(shaders, "froxelVolumeMaterial.comp", set())
    structs = {m.group(1): conv.parse_members(m.group(2)) for m in re.finditer(r"\bstruct\s+(\w+)\s*\{(.*?)\}\s*;", text, flags=re.S)}
    consts = dict(re.findall(r"^\s*const\s+(?:int|uint)\s+(\w+)\s*=\s*(\d+)\s*;", text, flags=re.M))
    block = re.search(r"uniform\s+global\s*\{(.*?)\}\s*;", text, flags=re.S).group(1)
    lay = conv.Layout(structs, consts, "std140")
    off, got = 0, {}
    for mt, mn, ml in conv.parse_members(block):
        a, s, _ = lay.member_info(mt, ml)
        off = conv.round_up(off, a)
        got[mn] = off
        off += s
    G = ffi.GlobalShaderInfo
    want = {"g_" + name: getattr(G, name).offset for name, _ in G._fields_}
    assert got == want and off == C.sizeof(G) == 340
    std430 = conv.Layout(structs, consts, "std430")
    assert std430.type_info("ShadowCascadeInfo")[1] == C.sizeof(ffi.ShadowCascadeInfo) == 304
    assert std430.type_info("LightBuffer")[1] in (20, 32)  # 20 bytes of members; the struct's own alignment rounds its array stride to 32
    assert conv.Layout(structs, consts, "std140").type_info("VolumetricLightingSettings")[2][-1][2] == 48  # phaseFunctionG, the 13th float
    assert std430.same_as_natural("BoundingBox") and std430.same_as_natural("CulledInstancesPerTile") and std430.same_as_natural("SDFInstance")
