"""temporalFilter.comp (TAA resolve, temporalFilter.comp:84-179 + temporalReprojection.inc) of the oracle against an independent float64
numpy restatement written from the GLSL: 3x3 neighbourhood with the reversible tonemap, resolve weights, closest-depth motion dilation,
bilinear history, AABB clip / clamp, contrast-driven blend factor, off-screen and camera-cut fallbacks. Compared to the precision of the
R11G11B10 targets."""
import numpy as np
import pytest

import passes
from conftest import decode_r11g11b10, random_r11g11b10

LUM = np.array([0.21, 0.72, 0.07])  # luminance.inc:5-7


def tonemap(c):
    return c / (1 + c @ LUM)[..., None]


def tonemap_reverse(c):
    return c / (1 - c @ LUM)[..., None]


def bilinear_clamp(img, u, v):
    h, w = img.shape[:2]
    fx, fy = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(fx).astype(int), np.floor(fy).astype(int)
    ax, ay = (fx - x0)[..., None], (fy - y0)[..., None]
    t = lambda xi, yi: img[np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)]
    return t(x0, y0) * (1 - ax) * (1 - ay) + t(x0 + 1, y0) * ax * (1 - ay) + t(x0, y0 + 1) * (1 - ax) * ay + t(x0 + 1, y0 + 1) * ax * ay


def np_taa(cur, his, motion, depth, weights, use_clipping, use_dilation, use_tonemap, camera_cut, history_tech=0):
    h, w = cur.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = (xs + 0.5) / w, (ys + 0.5) / h
    tm = tonemap if use_tonemap else (lambda c: c)

    def neighbourhood(img, uu, vv):
        return {(x, y): tm(bilinear_clamp(img, uu + x / w, vv + y / h)) for x in (-1, 0, 1) for y in (-1, 0, 1)}
    nb = neighbourhood(cur, u, v)
    stack = np.stack(list(nb.values()))
    mn, mx = stack.min(0), stack.max(0)
    current = sum(nb[(x, y)] * weights[(y + 1) * 3 + (x + 1)] for x in (-1, 0, 1) for y in (-1, 0, 1))  # w{x}_{y}, buffer order x fastest
    if use_dilation:  # getClosestFragmentMotion: strictly greater depth wins, x outer / y inner loop, texelFetch outside the image = 0
        best = np.zeros((h, w))
        off = np.zeros((h, w, 2), int)
        pad = np.pad(depth, 1)
        for x in (-1, 0, 1):
            for y in (-1, 0, 1):
                d = pad[1 + y:1 + y + h, 1 + x:1 + x + w]
                closer = d > best
                best = np.where(closer, d, best)
                off[closer] = (x, y)
        sx, sy = xs + off[..., 0], ys + off[..., 1]
        inside = (sx >= 0) & (sx < w) & (sy >= 0) & (sy < h)
        mot = np.where(inside[..., None], motion[np.clip(sy, 0, h - 1), np.clip(sx, 0, w - 1)], 0.0)
    else:
        mot = motion
    ur, vr = u + mot[..., 0], v + mot[..., 1]
    if history_tech == 0:
        hist = bilinear_clamp(his, ur, vr)
    elif history_tech == 1:  # bicubicSample16Tap, bicubicSampling.inc:28-70: Catmull-Rom weights at distances |d| + 1, |d|, 1 - |d|, 2 - |d|
        px, py = xs + 0.5 + mot[..., 0] * w, ys + 0.5 + mot[..., 1] * h

        def cr(d):   # catmullRomWeight1D for d >= 0 (every argument above is: the signed `d` of its second branch is not observable)
            return np.where(d <= 1, (9 * d ** 3 - 15 * d ** 2 + 6) / 6, np.where(d <= 2, (-3 * d ** 3 + 15 * d ** 2 - 24 * d + 12) / 6, 0.0))
        tx, ty = np.floor(px - 0.5) + 0.5, np.floor(py - 0.5) + 0.5
        dx, dy = np.abs(px - tx), np.abs(py - ty)
        wx, wy = [cr(dx + 1), cr(dx), cr(1 - dx), cr(2 - dx)], [cr(dy + 1), cr(dy), cr(1 - dy), cr(2 - dy)]
        hist = sum(bilinear_clamp(his, (tx + i - 1) / w, (ty + j - 1) / h) * (wx[i] * wy[j])[..., None] for j in range(4) for i in range(4))
    elif history_tech in (2, 3):  # bicubicSample9Tap :75-110 / bicubicSample5Tap :115-145: the two middle taps merged into one bilinear tap
        px, py = xs + 0.5 + mot[..., 0] * w, ys + 0.5 + mot[..., 1] * h

        def taps_1d(p, n):
            trunc = np.floor(p - 0.5) + 0.5
            f = p - trunc
            w0, w1 = -0.5 * f ** 3 + f ** 2 - 0.5 * f, 1.5 * f ** 3 - 2.5 * f ** 2 + 1
            w2, w3 = -1.5 * f ** 3 + 2 * f ** 2 + 0.5 * f, 0.5 * f ** 3 - 0.5 * f ** 2
            return [((trunc - 1) / n, w0), ((trunc + w2 / (w1 + w2)) / n, w1 + w2), ((trunc + 2) / n, w3)]
        X, Y = taps_1d(px, w), taps_1d(py, h)
        pairs = [(i, j) for i in range(3) for j in range(3) if history_tech == 2 or i == 1 or j == 1]   # the 5-tap variant drops the corners
        acc = sum(bilinear_clamp(his, X[i][0], Y[j][0]) * (X[i][1] * Y[j][1])[..., None] for i, j in pairs)
        hist = acc if history_tech == 2 else acc / sum(X[i][1] * Y[j][1] for i, j in pairs)[..., None]   # ... and renormalises
    else:  # bicubicSample1Tap (the default), bicubicSampling.inc:150-181: one bilinear tap + the current frame's cross neighbourhood
        px, py = xs + 0.5 + mot[..., 0] * w, ys + 0.5 + mot[..., 1] * h

        def weights_1d(p):
            trunc = np.floor(p - 0.5) + 0.5
            f = p - trunc
            w0, w1 = -0.5 * f ** 3 + f ** 2 - 0.5 * f, 1.5 * f ** 3 - 2.5 * f ** 2 + 1
            w2, w3 = -1.5 * f ** 3 + 2 * f ** 2 + 0.5 * f, 0.5 * f ** 3 - 0.5 * f ** 2
            return trunc, w0, w1 + w2, w3, w2 / (w1 + w2)
        tx, w0x, wBx, w3x, t_x = weights_1d(px)
        ty, w0y, wBy, w3y, t_y = weights_1d(py)
        hs = bilinear_clamp(his, (tx + t_x) / w, (ty + t_y) / h)
        c = nb[(0, 0)]
        terms = [(hs + nb[(-1, 0)] - c, w0x * wBy), (hs + nb[(0, -1)] - c, wBx * w0y), (hs, wBx * wBy), (hs + nb[(0, 1)] - c, wBx * w3y), (hs + nb[(1, 0)] - c, w3x * wBy)]
        hist = sum(v * wt[..., None] for v, wt in terms) / sum(wt for _, wt in terms)[..., None]
    hist = tm(hist)
    if use_clipping:  # clipAABB, temporalReprojection.inc:8-30
        centre, extend = 0.5 * (mx + mn), 0.5 * (mx - mn) + 0.0001
        to = hist - centre
        m = np.abs(to / extend).max(-1, keepdims=True)
        hist = np.where(m < 1, hist, centre + to / np.maximum(m, 1e-30))
    else:
        hist = np.clip(hist, mn, mx)
    lum = lambda c: c @ LUM
    contrast = lambda n: sum(np.abs(lum(n[k]) - lum(n[(0, 0)])) for k in n if k != (0, 0))
    last = neighbourhood(his, ur, vr)
    change = np.clip(np.abs(contrast(nb) - contrast(last)), 0, 1)
    blend = 0.13 * (1 - change) + 0.03 * change
    if camera_cut:
        blend = np.ones_like(blend)
    off_screen = (ur < 0) | (vr < 0) | (ur > 1) | (vr > 1)
    gauss = (nb[(-1, -1)] + nb[(-1, 1)] + nb[(1, -1)] + nb[(1, 1)]) * 0.0625 + (nb[(0, -1)] + nb[(-1, 0)] + nb[(0, 1)] + nb[(1, 0)]) * 0.125 + nb[(0, 0)] * 0.25
    blend = np.where(off_screen, 1.0, blend)
    current = np.where(off_screen[..., None], gauss, current)
    color = hist * (1 - blend[..., None]) + current * blend[..., None]
    return tonemap_reverse(color) if use_tonemap else color


@pytest.mark.parametrize("use_clipping,use_dilation,use_tonemap,camera_cut,history_tech", [(True, True, True, False, 4), (True, True, True, False, 0), (False, True, True, False, 0),
                                                                                         (True, False, False, False, 4), (True, True, True, True, 0),
                                                                                         (True, True, True, False, 1), (True, True, True, False, 2), (True, True, True, False, 3)])
def test_taa_resolve_matches_float64_restatement(ffi, oracle, use_clipping, use_dilation, use_tonemap, camera_cut, history_tech):
    rng = np.random.default_rng(40 + use_clipping * 2 + use_dilation)
    w, h = 48, 36
    # smooth HDR images (a TAA input is spatially coherent): low-frequency colour + mild noise, values 0.05 .. 8
    ys, xs = np.mgrid[0:h, 0:w]
    base = np.stack([1.5 + np.sin(xs / 7.0 + k) * np.cos(ys / 5.0 - k) for k in range(3)], -1) * np.array([2.0, 1.0, 0.5])

    def pack(img):  # quantise to R11G11B10 through the packed random generator's codec: encode by nearest representable value
        from conftest import decode_small_float
        out = np.zeros(img.shape[:2], np.uint32)
        for c, (mbits, shift) in enumerate(((6, 0), (6, 11), (5, 22))):
            codes = np.arange(1 << (5 + mbits))
            vals = decode_small_float(codes, mbits)
            vals[~np.isfinite(vals)] = np.inf
            idx = np.abs(img[..., c][..., None] - vals[None, None, :]).argmin(-1)
            out |= idx.astype(np.uint32) << shift
        return out
    cur_p = pack(base * rng.uniform(0.9, 1.1, (h, w, 3)))
    his_p = pack(base * rng.uniform(0.8, 1.2, (h, w, 3)) * 1.1)
    cur, his = decode_r11g11b10(cur_p), decode_r11g11b10(his_p)
    motion_i = rng.integers(-1500, 1500, (h, w, 2)).astype(np.int16)   # up to ~2 pixels
    motion_i[5:9, 10:20] = 20000                                        # a patch whose history lies off screen
    motion = np.maximum(motion_i.astype(np.float64) / 32767, -1)
    depth = rng.uniform(0.001, 0.05, (h, w)).astype(np.float32)
    jitter = rng.uniform(-0.5, 0.5, 2)
    wts = np.array([np.exp(-2.29 * ((jitter[0] - x) ** 2 + (jitter[1] - y) ** 2)) for y in (-1, 0, 1) for x in (-1, 0, 1)])
    wts /= wts.sum()
    out_p, hist_p = passes.taa_resolve(ffi, oracle, cur_p, his_p, motion_i, depth, wts, use_clipping, use_dilation, history_tech, use_tonemap, camera_cut)
    assert np.array_equal(out_p, hist_p)  # both targets receive the same colour (temporalFilter.comp:177-178)
    got = decode_r11g11b10(out_p)
    want = np_taa(cur, his, motion, depth.astype(np.float64), wts.astype(np.float32).astype(np.float64), use_clipping, use_dilation, use_tonemap, camera_cut, history_tech)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    # the blend-factor contrast term and the clip decision (m < 1) are discontinuous: allow a handful of pixels to land on the other side
    ok = (rel[..., :2].max(-1) < 2.0 ** -6 * 1.3) & (rel[..., 2] < 2.0 ** -5 * 1.3)
    assert ok.mean() > 0.995, "%d of %d pixels differ (max rel %.3f)" % (int((~ok).sum()), ok.size, rel.max())
    assert np.median(rel) < 2.0 ** -7
