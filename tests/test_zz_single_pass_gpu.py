"""Single passes of the GI denoiser on adversarial inputs (NaN texels, fast / off-screen motion, camera cut, depth edges), CUDA against
the oracle through the C-ABI, bit-exact - the inputs of tests/test_gi_temporal_upscale_numpy.py, which pins the oracle side.
First seen green on a B200 in round 2 (profiles/r2a_staged_tests.md)."""
import numpy as np
import pytest

import passes
from test_gi_temporal_upscale_numpy import gi_inputs

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("w,h,max_px,cut", [(48, 30, 2.0, False), (37, 23, 6.0, False), (40, 24, 0.0, False), (32, 20, 2.0, True), (130, 70, 3.0, False)])
def test_gi_temporal_filter_bit_exact(ffi, cuda, oracle, w, h, max_px, cut):
    rng = np.random.default_rng(w * 100 + h)
    y_sh, co_cg, hist_y, hist_c, cur, last = gi_inputs(rng, w, h, max_px)
    y_sh[3, 5, 1] = np.nan
    hist_y[7, 2, :] = np.nan
    co_cg[9, 9, 0] = np.nan
    hist_c[9, 9, 1] = np.nan
    a = passes.gi_temporal_filter(ffi, cuda, y_sh, co_cg, hist_y, hist_c, cur, last, camera_cut=cut)
    b = passes.gi_temporal_filter(ffi, oracle, y_sh, co_cg, hist_y, hist_c, cur, last, camera_cut=cut)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint16), y.view(np.uint16))


@pytest.mark.parametrize("W,H", [(64, 40), (50, 30), (33, 21), (258, 130)])
def test_gi_upscale_bit_exact(ffi, cuda, oracle, W, H):
    rng = np.random.default_rng(W + H)
    w, h = W // 2, H // 2
    ys, xs = np.mgrid[0:H, 0:W]
    lin = np.where(xs + 0.6 * ys > 0.55 * W, 9.0, 2.0) + 0.05 * rng.random((H, W))
    depth_full = (1 - (0.1 * 300.0 / lin - 300.0) / (0.1 - 300.0)).astype(np.float32)
    depth_full[0:3, 0:4] = 0.0
    depth_half = depth_full[::2, ::2][:h, :w].astype(np.float16)
    y_sh = (rng.random((h, w, 4)) * 2 - 0.5).astype(np.float16)
    co_cg = (rng.random((h, w, 2)) - 0.5).astype(np.float16)
    y_sh[2, 3, 0] = np.nan
    a, b = passes.gi_upscale(ffi, cuda, y_sh, co_cg, depth_full, depth_half), passes.gi_upscale(ffi, oracle, y_sh, co_cg, depth_full, depth_half)
    assert np.array_equal(a[0].view(np.uint16), b[0].view(np.uint16)) and np.array_equal(a[1].view(np.uint16), b[1].view(np.uint16))


@pytest.mark.parametrize("res,moving,cut", [((12, 7, 16), False, False), ((10, 6, 8), True, False), ((9, 5, 8), True, True), ((60, 34, 64), True, False), ((19, 9, 70), True, False)])
def test_froxel_passes_bit_exact(ffi, cuda, oracle, res, moving, cut):
    """the four froxel passes launched one by one (all four volumes compared) and as the fused column launch, the product's default (the
    reprojected and the integrated volume compared; the launch leaves the material and scattering volumes alone)"""
    from test_froxels_numpy import scene
    cam, prev, noise, shadow, L, settings, light, history, sun = scene(res[0] * 10 + res[2], res, moving)
    history[1, 2, 3, :] = np.nan
    a = passes.froxels(ffi, cuda, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    b = passes.froxels(ffi, oracle, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    for name, x, y in zip(("material", "scattering", "reprojection", "integration"), a, b):
        assert np.array_equal(x.view(np.uint16), y.view(np.uint16)), name
    fused = passes.froxels(ffi, cuda, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut, pass_fusion=True)
    for name, x, y in list(zip(("material", "scattering", "reprojection", "integration"), fused, b))[2:]:
        assert np.array_equal(x.view(np.uint16), y.view(np.uint16)), name + " (fused column launch)"
    assert not fused[0].view(np.uint16).any() and not fused[1].view(np.uint16).any(), "the fused launch wrote the material / scattering volume"


@pytest.mark.parametrize("w,h,radius,strength", [(64, 48, 1.5, 0.05), (100, 75, 1.0, 0.3), (33, 17, 2.5, 1.0)])
def test_bloom_with_hot_texels_bit_exact(ffi, cuda, oracle, w, h, radius, strength):
    rng = np.random.default_rng(w + h)
    from conftest import random_r11g11b10
    packed = random_r11g11b10(rng, w * h, finite=False).reshape(h, w)   # inf / NaN texels included
    a, b = passes.bloom(ffi, cuda, packed, strength=strength, radius=radius), passes.bloom(ffi, oracle, packed, strength=strength, radius=radius)
    for x, y in zip(a[0] + a[1] + [a[2]], b[0] + b[1] + [b[2]]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("case", ["bright", "single_bin", "adapting_up"])
def test_pre_expose_lights_bit_exact(ffi, cuda, oracle, case):
    from conftest import random_r11g11b10
    rng = np.random.default_rng(3)
    screen = (320, 180)
    hist = np.zeros(128, np.uint32)
    if case == "single_bin":
        hist[80] = screen[0] * screen[1]
    else:
        hist = np.bincount(np.clip(rng.normal(95, 6, screen[0] * screen[1]).round(), 0, 127).astype(int), minlength=128).astype(np.uint32)
    light = np.array([1, 1, 1, 1e-6 if case == "adapting_up" else 2e-5, 1], np.float32)
    lut = random_r11g11b10(rng, 16 * 24).reshape(24, 16)
    a, b = passes.pre_expose_lights(ffi, cuda, hist, light, lut, -0.61, screen), passes.pre_expose_lights(ffi, oracle, hist, light, lut, -0.61, screen)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("sun,cascades", [((0.35, -0.8, 0.48), 4), ((0.0, -1.0, 0.0), 4), ((-0.6, -0.3, -0.74), 3)])
def test_light_matrix_bit_exact(ffi, cuda, oracle, sun, cascades):
    sun = np.array(sun) / np.linalg.norm(sun)
    fwd = np.array([0.6, 0.1, -0.79])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0, -1.0, 0])
    right /= np.linalg.norm(right)
    cam = dict(position=np.array([3.0, -2.0, 1.0]), forward=fwd, up=np.cross(right, fwd), right=right, tan_fov_half=0.41, aspect=16 / 9)
    dmm = np.array([0.0021, 0.083], np.float32)
    a, b = passes.light_matrix(ffi, cuda, dmm, cam, sun, cascades=cascades), passes.light_matrix(ffi, oracle, dmm, cam, sun, cascades=cascades)
    assert bytes(a) == bytes(b)


@pytest.mark.parametrize("sun", [(0.3, -0.5, 0.81), (0.0, -1.0, 0.0), (0.7, 0.05, 0.7)])
def test_sky_luts_bit_exact(ffi, cuda, oracle, sun):
    sun = np.array(sun) / np.linalg.norm(sun)
    for x, y in zip(passes.sky_luts(ffi, cuda, sun, 3.0), passes.sky_luts(ffi, oracle, sun, 3.0)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("strict,with_box", [(False, False), (False, True), (True, True)])
def test_sdf_diffuse_trace_single_pass_bit_exact(ffi, cuda, oracle, strict, with_box):
    from test_sdf_diffuse_trace_numpy import wall_inputs
    from test_sdf_trace_analytic import box_brick
    rng = np.random.default_rng(6)
    depth, normal, noise, sky_p = wall_inputs(rng, 64, 40)
    half, centre = np.array([1.6, 1.2, 0.9]), np.array([0.3, 0.2, -6.5])
    pad = np.maximum(2 * half * 0.075, 0.5)
    W2L = np.eye(4)
    W2L[:3, 3] = -centre
    Lm = np.array([[1 / 20.0, 0, 0, 0], [0, 0, 1 / 20.0, 0], [0, 1 / 40.0, 0, 0.5], [0, 0, 0, 1]], np.float64)
    shadow = np.zeros((16, 16), np.uint16)
    shadow[:, 8:] = 65535
    inst = [(2 * (half + pad), 0, np.array([0.7, 0.5, 0.3]), W2L)] if with_box else []
    bricks = [box_brick(half, 32).reshape(32, 32, 32)] if with_box else []
    a = passes.sdf_diffuse_trace(ffi, cuda, depth, normal, noise, sky_p, inst, bricks, shadow, Lm.T.ravel(), [1.0, 0.9, 0.8, 1.0, 2.0], 3.4, strict)
    b = passes.sdf_diffuse_trace(ffi, oracle, depth, normal, noise, sky_p, inst, bricks, shadow, Lm.T.ravel(), [1.0, 0.9, 0.8, 1.0, 2.0], 3.4, strict)
    assert np.array_equal(a[0].view(np.uint16), b[0].view(np.uint16)) and np.array_equal(a[1].view(np.uint16), b[1].view(np.uint16))


@pytest.mark.parametrize("n,use_hiz,influence", [(40, False, 5.0), (300, True, 2.0), (260, False, 40.0), (1200, True, 5.0)])
def test_sdf_culling_bit_exact(ffi, cuda, oracle, n, use_hiz, influence):
    from test_sdf_culling_numpy import camera_and_frustum
    rng = np.random.default_rng(n + use_hiz)
    cam, pts, nrm = camera_and_frustum()
    centres = cam["position"] + rng.uniform(-60, 60, (n, 3)) * np.array([1.0, 0.3, 1.0])
    halfs = rng.uniform(0.3, 4.0, (n, 3))
    bbs = np.stack([centres - halfs, centres + halfs], 1)
    hiz = rng.uniform(0.001, 0.09, (3, 5, 2)).astype(np.float32) if use_hiz else None
    if hiz is not None:
        hiz.sort(-1)
    a = passes.sdf_culling(ffi, cuda, bbs, pts, nrm, influence, (160, 96), (320, 192), cam, hiz)
    b = passes.sdf_culling(ffi, oracle, bbs, pts, nrm, influence, (160, 96), (320, 192), cam, hiz)
    assert np.array_equal(a[0], b[0])
    # entries beyond a tile's count are never written by either side
    for ta, tb in zip(a[1], b[1]):
        assert ta[0] == tb[0] and (ta[0] == 0xFFFFFFFF or np.array_equal(ta[1:1 + ta[0]], tb[1:1 + tb[0]]))


def _taa_inputs(rng, w, h, max_px, special):
    from conftest import random_r11g11b10
    cur = random_r11g11b10(rng, w * h, finite=True).reshape(h, w)
    his = random_r11g11b10(rng, w * h, finite=True).reshape(h, w)
    # keep most values moderate so that the resolve is not dominated by 2^15-sized texels
    cur = np.where(rng.random((h, w)) < 0.9, cur & np.uint32(0xBBFEFBFF), cur).astype(np.uint32)
    his = np.where(rng.random((h, w)) < 0.9, his & np.uint32(0xBBFEFBFF), his).astype(np.uint32)
    if special:  # a few inf / NaN texels: their blocks take the spelled-out path, the others the fast one
        for img in (cur, his):
            n = max(w * h // 700, 2)
            ys, xs = rng.integers(0, h, n), rng.integers(0, w, n)
            img[ys, xs] = random_r11g11b10(rng, n, finite=False) | np.uint32(0x7C0)
    motion = np.zeros((h, w, 2), np.int16)
    motion[..., 0] = np.clip(rng.normal(0, max_px, (h, w)) / w * 32767, -32767, 32767).astype(np.int16)
    motion[..., 1] = np.clip(rng.normal(0, max_px, (h, w)) / h * 32767, -32767, 32767).astype(np.int16)
    depth = rng.random((h, w)).astype(np.float32)
    depth[rng.random((h, w)) < 0.1] = 0.0
    wts = rng.random(9).astype(np.float32)
    wts /= wts.sum()
    return cur, his, motion, depth, wts


@pytest.mark.parametrize("w,h,max_px,special,tech,clip,dil,tonemap", [
    (200, 120, 0.7, False, 4, True, True, True),     # the default configuration: interior blocks on the fast path, a border ring on the generic one
    (200, 120, 9.0, False, 4, True, True, True),     # motion beyond the staged history tile: taps fall back to global loads
    (200, 120, 1.5, True, 4, True, True, True),      # inf / NaN texels
    (131, 77, 2.0, True, 0, False, False, False),    # odd extent, bilinear history, clamp instead of clip, no dilation, no tonemap
    (160, 96, 2.0, False, 1, True, True, True), (160, 96, 2.0, True, 2, True, False, True), (160, 96, 3.0, False, 3, False, True, False)])
def test_taa_resolve_single_pass_bit_exact(ffi, cuda, oracle, w, h, max_px, special, tech, clip, dil, tonemap):
    rng = np.random.default_rng(w * 7 + h + tech)
    cur, his, motion, depth, wts = _taa_inputs(rng, w, h, max_px, special)
    a = passes.taa_resolve(ffi, cuda, cur, his, motion, depth, wts, clip, dil, tech, tonemap)
    b = passes.taa_resolve(ffi, oracle, cur, his, motion, depth, wts, clip, dil, tech, tonemap)
    for name, x, y in zip(("output", "history"), a, b):
        assert np.array_equal(x, y), "%s: %d of %d texels differ" % (name, int((x != y).sum()), x.size)


@pytest.mark.parametrize("w,h,radius,finite", [(512, 288, 1.5, True), (480, 270, 1.5, False), (258, 130, 0.75, True), (300, 200, 100.0, True)])
def test_bloom_chain_interior_blocks_bit_exact(ffi, cuda, oracle, w, h, radius, finite):
    """Sizes with interior blocks in several mips (the fast path of the bloom kernels), odd mip extents, a blur radius beyond the fast path's bound."""
    rng = np.random.default_rng(w + h)
    from conftest import random_r11g11b10
    packed = random_r11g11b10(rng, w * h, finite=finite).reshape(h, w)
    a, b = passes.bloom(ffi, cuda, packed, strength=0.05, radius=radius), passes.bloom(ffi, oracle, packed, strength=0.05, radius=radius)
    for x, y in zip(a[0] + a[1] + [a[2]], b[0] + b[1] + [b[2]]):
        assert np.array_equal(x, y)
