"""filterIndirectDiffuseTemporal.comp (S5) and indirectLightUpscale.comp (S6) of the oracle against independent float64 numpy
restatements written from the GLSL: bilinear history reprojection along the motion vector, the SMAA-style motion-difference blend
factor, the fast-motion / off-screen / camera-cut / NaN paths; the depth-aware upscale with textureGather's texel order, the
closest-depth texel on edges and bilinear elsewhere. Results are half floats: agreement to half precision (the oracle computes in
binary32), apart from texels whose reprojected position sits on a texel border or whose blend factor sits on a branch threshold."""
import numpy as np
import pytest

import passes


def bilinear(img, u, v, repeat=False):
    """VK linear filter, clamp-to-edge (or repeat) addressing: taps around u * size - 0.5, weights = the fractions."""
    h, w = img.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    def idx(i, n):
        i = i.astype(np.int64)
        return np.mod(i, n) if repeat else np.clip(i, 0, n - 1)
    xa, xb, ya, yb = idx(x0, w), idx(x0 + 1, w), idx(y0, h), idx(y0 + 1, h)
    img = img.astype(np.float64).reshape(h, w, -1)
    return (img[ya, xa] * (1 - fx) + img[ya, xb] * fx) * (1 - fy) + (img[yb, xa] * (1 - fx) + img[yb, xb] * fx) * fy


def nearest(img, u, v):
    h, w = img.shape[:2]
    return img[np.clip(np.floor(v * h).astype(np.int64), 0, h - 1), np.clip(np.floor(u * w).astype(np.int64), 0, w - 1)]


def snorm16(m):
    return np.maximum(m.astype(np.float64) / 32767.0, -1.0)


def np_temporal(y_sh, co_cg, hist_y, hist_c, motion_cur, motion_last, camera_cut):  # filterIndirectDiffuseTemporal.comp:20-86
    h, w = y_sh.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = (xs + 0.5) / w, (ys + 0.5) / h
    cur_y, cur_c = bilinear(y_sh, u, v), bilinear(co_cg, u, v)
    motion = bilinear(snorm16(motion_cur), u, v)
    ur, vr = u + motion[..., 0], v + motion[..., 1]
    his_y, his_c = bilinear(hist_y, ur, vr), bilinear(hist_c, ur, vr)
    motion_prev = bilinear(snorm16(motion_last), ur, vr, repeat=True)
    length = lambda a: np.sqrt((a * a).sum(-1))
    diff = np.sqrt(np.abs(length(motion) - length(motion_prev)))
    factor = np.clip(diff * 10, 0, 1)
    alpha_min = np.maximum(0.6 - 0.3 * np.abs(length(cur_y) - length(his_y)), 0)
    alpha = 0.8 * (1 - factor) + alpha_min * factor
    res = np.array([2 * w, 2 * h], np.float64)  # g_screenResolution: the full resolution
    fast = (np.abs(motion) * res > 3).any(-1) | (np.abs(motion_prev) * res > 3).any(-1)
    alpha = np.where(fast, alpha_min, alpha)
    alpha = np.where((ur < 0) | (vr < 0) | (ur > 1) | (vr > 1), 0.0, alpha)
    if camera_cut:
        alpha = np.zeros_like(alpha)
    nan_cur = np.isnan(cur_y).any(-1) | np.isnan(cur_c).any(-1)
    alpha = np.where(nan_cur, 1.0, alpha)
    his_y = np.where((nan_cur & np.isnan(his_y).any(-1))[..., None], 0.0, his_y)
    his_c = np.where((nan_cur & np.isnan(his_c).any(-1))[..., None], 0.0, his_c)
    a = alpha[..., None]
    with np.errstate(invalid="ignore"):
        return cur_y * (1 - a) + his_y * a, cur_c * (1 - a) + his_c * a, alpha, (ur, vr)


def gi_inputs(rng, w, h, max_motion_px):
    y_sh = (rng.random((h, w, 4)) * 2 - 0.5).astype(np.float16)
    co_cg = (rng.random((h, w, 2)) - 0.5).astype(np.float16)
    hist_y = (y_sh.astype(np.float32) + rng.normal(0, 0.3, (h, w, 4))).astype(np.float16)
    hist_c = (rng.random((h, w, 2)) - 0.5).astype(np.float16)
    # smooth motion field (full resolution, uv units) with a few fast regions; last frame's differs slightly
    ys, xs = np.mgrid[0:2 * h, 0:2 * w]
    base = np.stack([np.sin(xs / 9.0) * np.cos(ys / 7.0), np.cos(xs / 5.0 + ys / 11.0)], -1) * max_motion_px / np.array([2 * w, 2 * h])
    cur = np.round(base * 32767).astype(np.int16)
    last = np.round((base * (1 + 0.4 * rng.random((2 * h, 2 * w, 1))) * 32767)).astype(np.int16)
    return y_sh, co_cg, hist_y, hist_c, cur, last


@pytest.mark.parametrize("w,h,max_px,cut", [(48, 30, 2.0, False), (37, 23, 6.0, False), (40, 24, 0.0, False), (32, 20, 2.0, True)])
def test_gi_temporal_filter_matches_numpy(ffi, oracle, w, h, max_px, cut):
    rng = np.random.default_rng(w * 100 + h)
    y_sh, co_cg, hist_y, hist_c, cur, last = gi_inputs(rng, w, h, max_px)
    if not cut:
        y_sh[3, 5, 1] = np.nan           # the NaN filter: alpha = 1, history kept
        hist_y[7, 2, :] = np.nan         # NaN history is only cleared when the current texel is NaN too: it propagates
    out_y, out_c, ho_y, ho_c = passes.gi_temporal_filter(ffi, oracle, y_sh, co_cg, hist_y, hist_c, cur, last, camera_cut=cut)
    assert np.array_equal(out_y.view(np.uint16), ho_y.view(np.uint16)) and np.array_equal(out_c.view(np.uint16), ho_c.view(np.uint16))  # :82-85 the same value twice
    ref_y, ref_c, alpha, (ur, vr) = np_temporal(y_sh, co_cg, hist_y, hist_c, cur, last, cut)
    finite = np.isfinite(ref_y).all(-1)
    assert np.array_equal(np.isfinite(out_y.astype(np.float64)).all(-1), finite)
    err_y = np.abs(out_y.astype(np.float64) - ref_y)[finite].max(-1)
    err_c = np.abs(out_c.astype(np.float64) - ref_c)[finite].max(-1)
    tol = 2.0 ** -10 * 2.5                                   # half an ulp of a half float below 4 + binary32 rounding of the taps
    # a reprojected position within rounding distance of an edge of the unit square flips the off-screen test
    near_edge = (np.minimum(np.abs(ur), np.abs(ur - 1)) < 1e-6) | (np.minimum(np.abs(vr), np.abs(vr - 1)) < 1e-6)
    bad = ((err_y > tol) | (err_c > tol)) & ~near_edge[finite]
    assert bad.mean() <= 0.002, "%d of %d texels differ, max %.3e" % (bad.sum(), bad.size, max(err_y.max(), err_c.max()))
    if cut:
        assert np.abs(out_y.astype(np.float64) - y_sh.astype(np.float64)).max() <= 2.0 ** -10  # alpha = 0: the current frame (texel centres: one tap)
    else:
        assert 0.05 < alpha[finite].mean() < 0.81                                    # 0.8 = alphaDefault where nothing moves
        assert (alpha == 0).any() == bool(((ur < 0) | (vr < 0) | (ur > 1) | (vr > 1)).any())
        assert max_px < 4 or (alpha[finite] < 0.6).any()                             # fast motion: alpha = alphaMin, lowered by the luminance difference


def linearize(d, near, far):  # linearDepth.inc:5-8
    return near * far / (far + (1 - d) * (near - far))


def np_upscale(y_sh, co_cg, depth_full, depth_half, g):  # indirectLightUpscale.comp:17-71
    H, W = depth_full.shape
    h, w = depth_half.shape
    ys, xs = np.mgrid[0:H, 0:W]
    u, v = (xs + 0.5) / g.screenResolution[0], (ys + 0.5) / g.screenResolution[1]
    d_full = linearize(nearest(depth_full, u, v).astype(np.float64), g.nearPlane, g.farPlane)
    # textureGather: the four texels a linear filter would blend, in the order (i0, j0+1), (i0+1, j0+1), (i0+1, j0), (i0, j0)
    i0, j0 = np.floor(u * w - 0.5).astype(np.int64), np.floor(v * h - 0.5).astype(np.int64)
    cx, cy = lambda i: np.clip(i, 0, w - 1), lambda j: np.clip(j, 0, h - 1)
    dh = depth_half.astype(np.float64)
    samples = [dh[cy(j0 + 1), cx(i0)], dh[cy(j0 + 1), cx(i0 + 1)], dh[cy(j0), cx(i0 + 1)], dh[cy(j0), cx(i0)]]
    offsets = [(0, 1), (1, 1), (1, 0), (0, 0)]
    min_diff = np.full((H, W), 1000.0)
    closest = np.zeros((H, W, 2))
    edge = np.zeros((H, W), bool)
    for s, o in zip(samples, offsets):
        diff = np.abs(linearize(s, g.nearPlane, g.farPlane) - d_full)
        edge |= diff > 0.5
        take = diff < min_diff
        min_diff = np.where(take, diff, min_diff)
        closest = np.where(take[..., None], np.array(o, np.float64), closest)
    uc, vc = u + closest[..., 0] / w, v + closest[..., 1] / h        # :58 offset from the FULL-res pixel centre, as the shader does
    e = edge[..., None]
    return (np.where(e, nearest(y_sh, uc, vc).astype(np.float64), bilinear(y_sh, u, v)), np.where(e, nearest(co_cg, uc, vc).astype(np.float64), bilinear(co_cg, u, v)), edge, min_diff)


@pytest.mark.parametrize("W,H", [(64, 40), (50, 30), (33, 21)])
def test_gi_upscale_matches_numpy(ffi, oracle, W, H):
    rng = np.random.default_rng(W + H)
    w, h = W // 2, H // 2
    near, far = 0.1, 300.0
    # two depth layers (2 m and 9 m) split by a slanted line: edges where the gathered half-res depths straddle the line
    ys, xs = np.mgrid[0:H, 0:W]
    lin = np.where(xs + 0.6 * ys > 0.55 * W, 9.0, 2.0) + 0.05 * rng.random((H, W))
    depth_full = (1 - (near * far / lin - far) / (near - far)).astype(np.float32)   # inverse of linearizeDepth
    depth_full[0:3, 0:4] = 0.0                                                       # sky
    depth_half = depth_full[::2, ::2][:h, :w].astype(np.float16)
    y_sh = (rng.random((h, w, 4)) * 2 - 0.5).astype(np.float16)
    co_cg = (rng.random((h, w, 2)) - 0.5).astype(np.float16)
    out_y, out_c, g = passes.gi_upscale(ffi, oracle, y_sh, co_cg, depth_full, depth_half)
    ref_y, ref_c, edge, min_diff = np_upscale(y_sh, co_cg, depth_full, depth_half, g)
    assert 0.03 < edge.mean() < 0.6
    err = np.maximum(np.abs(out_y.astype(np.float64) - ref_y).max(-1), np.abs(out_c.astype(np.float64) - ref_c).max(-1))
    # R16F depth near the 0.5 m threshold, or two candidates at (nearly) the same distance, may resolve differently in binary32
    assert (err > 2.0 ** -10 * 2.5).mean() <= 0.003, "%d texels differ, max %.3e" % ((err > 2.0 ** -10 * 2.5).sum(), err.max())
    # non-edge texels are the plain bilinear upscale, edge texels are copies of one half-res texel
    e = edge & (err == 0)
    assert e.any() and all((out_y[y, x] == y_sh.reshape(-1, 4)).all(-1).any() for y, x in zip(*np.nonzero(e)))
