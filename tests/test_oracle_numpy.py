"""The CPU oracle against independent numpy restatements of the exactly specified pieces of the frame path (the
reference ships no golden vectors, SURVEY.md 4 / 8c: this is how the oracle is pinned besides tests/golden)."""
import numpy as np
import pytest

from conftest import decode_r11g11b10, random_r11g11b10
import passes


def np_hiz(depth):
    """min (sky-excluded) / max pyramid, every level from the complete previous level, odd extents take the extra
    row/column (depthHiZPyramid.comp:52-124) with the '*' quirk on the odd x odd corner tap (:114)."""
    h, w = depth.shape
    levels = []
    src = None
    sw, sh = w, h
    first = True
    while True:
        dw, dh = max(sw // 2, 1), max(sh // 2, 1)
        out = np.zeros((dh, dw, 2), np.float32)
        for y in range(dh):
            for x in range(dw):
                taps = [(0, 0), (1, 0), (0, 1), (1, 1)]
                if sh % 2 == 1:
                    taps += [(0, 2), (1, 2)]
                if sw % 2 == 1:
                    taps += [(2, 0), (2, 1)]
                corner = sh % 2 == 1 and sw % 2 == 1
                if corner:
                    taps += [(2, 2)]
                mn, mx = np.float32(1.0), np.float32(0.0)
                for i, (ox, oy) in enumerate(taps):
                    sx, sy = min(2 * x + ox, sw - 1), min(2 * y + oy, sh - 1)
                    if first:
                        d = depth[sy, sx]
                        sky = np.float32(1.0 if d == 0 else 0.0)
                        v = d * sky if (corner and i == len(taps) - 1) else d + sky
                        mn, mx = min(mn, v), max(mx, d)
                    else:
                        t = src[sy, sx]
                        mn, mx = min(mn, t[0] + np.float32(1.0 if t[1] == 0 else 0.0)), max(mx, t[1])
                out[y, x] = (mn, mx)
        levels.append(out)
        if dw == 1 and dh == 1:
            break
        src, sw, sh, first = out, dw, dh, False
    return levels


@pytest.mark.parametrize("w,h", [(64, 36), (50, 30), (37, 23), (130, 66), (2, 2), (18, 5), (4096, 4), (4098, 6)])  # the last two: 12 levels
def test_hiz_matches_numpy(ffi, oracle, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    depth = rng.uniform(0.0005, 0.9, (h, w)).astype(np.float32)
    depth[rng.uniform(size=(h, w)) < 0.2] = 0.0  # sky
    got = passes.hiz(ffi, oracle, depth)
    want = np_hiz(depth)
    assert len(got) == len(want)
    for lvl, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "level %d" % lvl


def test_histogram_counts_and_bins(ffi, oracle):
    rng = np.random.default_rng(3)
    w, h = 96, 70  # partial tile row at the bottom (6 rows >= the 4 rows of invocations that write the 128 bins)
    packed = random_r11g11b10(rng, w * h).reshape(h, w)
    packed[:4] = 0  # black rows -> log(0) = -inf -> bin 0
    exposure = 0.37
    per_tile, hist = passes.histogram(ffi, oracle, packed, exposure)
    assert hist.sum() == w * h
    assert np.array_equal(per_tile.sum(axis=(0, 1)), hist)
    rgb = decode_r11g11b10(packed)
    lum = (rgb @ np.array([0.2126, 0.7152, 0.0722])) / exposure
    with np.errstate(divide="ignore"):
        t = (np.log(lum) - np.log(0.001)) / (np.log(200000.0) - np.log(0.001))
    bins = np.floor(127 * np.clip(t, 0, 1)).astype(np.int64)
    want = np.bincount(bins.ravel(), minlength=128)
    # float32 log vs float64 log: only pixels within rounding distance of a bin edge may move to the neighbouring bin
    assert np.abs(np.cumsum(hist.astype(np.int64)) - np.cumsum(want)).max() <= 3
    assert hist[0] >= 4 * w


def test_histogram_tile_row_quirk(ffi, oracle):
    """A last tile row with fewer than 4 pixel rows: bins >= 32*rows are never written by the reference (their
    invocations left the image, histogramPerTile.comp:37-39) - the per-tile buffer keeps its previous (zero) content."""
    rng = np.random.default_rng(5)
    w, h = 64, 34
    packed = random_r11g11b10(rng, w * h).reshape(h, w)
    per_tile, hist = passes.histogram(ffi, oracle, packed, 1.0)
    assert per_tile[1, :, 64:].sum() == 0
    assert hist.sum() == per_tile.sum() <= w * h
    # same for a last tile column narrower than 32 pixels: only bins b with b % 32 < columns are written
    w2, h2 = 100, 64
    per_tile2, hist2 = passes.histogram(ffi, oracle, random_r11g11b10(rng, w2 * h2).reshape(h2, w2), 1.0)
    mask = (np.arange(128) % 32) >= 4
    assert per_tile2[:, 3, mask].sum() == 0 and hist2.sum() < w2 * h2


def np_hash32(qx, qy):
    qx, qy = qx.astype(np.uint32), qy.astype(np.uint32)
    UI0, UI1, UI2 = np.uint32(1597334673), np.uint32(3812015801), np.uint32(2798796415)
    with np.errstate(over="ignore"):
        h = (qx * UI0) ^ (qy * UI1) ^ (qx * UI2)
        n = np.stack([h * UI0, h * UI1, h * UI2], axis=-1)
    return n.astype(np.float32) * np.float32(1.0 / np.float32(0xFFFFFFFF))


def test_tonemap_matches_numpy(ffi, oracle):
    rng = np.random.default_rng(11)
    w, h, time = 96, 40, 0.37
    packed = random_r11g11b10(rng, w * h).reshape(h, w)
    got = passes.tonemap(ffi, oracle, packed, time)
    c = np.minimum(decode_r11g11b10(packed), 65504.0)
    m_in = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]])
    m_out = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]])
    v = c @ m_in.T
    v = (v * (v + 0.0245786) - 0.000090537) / (v * (0.983729 * v + 0.4329510) + 0.238081)
    v = np.clip(v @ m_out.T, 0, 1)
    srgb = np.where(v <= 0.0031308, v * 12.92, 1.055 * np.power(v, 1 / 2.4) - 0.055)
    ys, xs = np.mgrid[0:h, 0:w]
    t = np.float32(time)
    f = lambda a: np.floor(a.astype(np.float32) * t)
    noise = np_hash32(f(xs), f(ys)).astype(np.float64) + np_hash32(np.floor((xs.astype(np.float32) + np.float32(165)) * t), np.floor((ys.astype(np.float32) + np.float32(1292)) * t))
    srgb = srgb + (noise - 1.0) / 255.0
    want = np.floor(np.clip(srgb, 0, 1) * 255 + 0.5).astype(np.int32)
    got_rgb = got[..., [2, 1, 0]].astype(np.int32)  # B8G8R8A8
    assert (got[..., 3] == 255).all()
    diff = np.abs(got_rgb - want)
    assert diff.max() <= 1  # float32 vs float64 may flip a rounding
    assert (diff == 0).mean() > 0.995


def test_depth_downscale_is_a_strided_copy(ffi, oracle):
    rng = np.random.default_rng(2)
    depth = rng.uniform(0, 1, (30, 44)).astype(np.float32)
    got = passes.depth_downscale(ffi, oracle, depth)
    assert np.array_equal(got, depth[::2, ::2][:15, :22].astype(np.float16))


def test_bloom_energy_and_mix(ffi, oracle):
    """A constant image stays constant through the 13-tap / tent chain (weights sum to 1) and through the mix."""
    w, h = 64, 48
    one = np.uint32((15 << 6) | ((15 << 6) << 11) | ((15 << 5) << 22))  # (1, 1, 1)
    packed = np.full((h, w), one, np.uint32)
    down, up, result = passes.bloom(ffi, oracle, packed)
    for m in down:
        assert (m == one).all()
    assert (up[-1] == one).all()  # the coarsest upsample is the 9-tap tent alone
    for level, m in enumerate(up):  # each finer level adds the 4-tap box of the coarser upsample mip: 1, 2, 3, 4, 5
        assert np.allclose(decode_r11g11b10(m), float(len(up) - level))
    assert np.allclose(decode_r11g11b10(result), 0.95 + 0.05 * 5.0, atol=1 / 64)


def test_halton_jitter_and_resolve_weights(ffi, oracle):
    """TAA jitter = Halton(2,3)[frame % 8] * 2 - 1 pixels (TAA.cpp:168-170), resolve weights exp(-2.29 d^2) normalised."""
    s = ffi.default_settings(oracle, 64, 36)
    fe = ffi.Frontend(oracle, s)
    cam = ffi.camera((0, -1.7, 0), (1, 0, 0), (0, 0, 1), (0, -1, 0))

    def radical(i, base):
        f, r = 1.0, 0.0
        while i:
            f /= base
            r += f * (i % base)
            i //= base
        return r
    for frame in range(1, 10):
        fe.render_frame(cam, frame / 60.0, 1 / 60.0)
        g = fe.global_shader_info()
        jx, jy = 2 * radical(frame % 8, 2) - 1, 2 * radical(frame % 8, 3) - 1
        assert g.currentFrameCameraJitter[0] == pytest.approx(jx / 64, abs=1e-7)
        assert g.currentFrameCameraJitter[1] == pytest.approx(jy / 36, abs=1e-7)
        wts = np.array([np.exp(-2.29 * ((jx - x) ** 2 + (jy - y) ** 2)) for y in (-1, 0, 1) for x in (-1, 0, 1)])
        assert np.allclose(fe.resolve_weights(), wts / wts.sum(), rtol=2e-5)
        assert g.frameIndexMod4 == frame % 4
    fe.close()


# ---------------- colorToLuminance.comp + temporalSupersampling.comp ----------------
def test_color_to_luminance_matches_numpy(ffi, oracle):
    rng = np.random.default_rng(21)
    w, h = 70, 37
    packed = random_r11g11b10(rng, w * h).reshape(h, w)
    packed[rng.uniform(size=(h, w)) < 0.5] &= np.uint32(0x0FF3FDFF)  # half of the texels below ~2 so that the R8 target is not saturated everywhere
    got = passes.color_to_luminance(ffi, oracle, packed)
    lum = decode_r11g11b10(packed) @ np.array([0.21, 0.72, 0.07])  # luminance.inc:5-7
    want = np.floor(np.clip(lum, 0, 1) * 255 + 0.5).astype(np.int32)
    diff = np.abs(got.astype(np.int32) - want)
    assert diff.max() <= 1 and (diff == 0).mean() > 0.99  # float32 vs float64 may flip a rounding
    assert 0.1 < (got == 255).mean() < 0.9


def pack_r11g11b10_exact(r, g, b):
    """pack values that are exactly representable (powers of two times small integers)"""
    def chan(v, mbits):
        m, e = np.frexp(np.float64(v))  # v = m * 2^e, m in [0.5, 1)
        return (int(e - 1 + 15) << mbits) | int(round((m * 2 - 1) * (1 << mbits)))
    return np.uint32(chan(r, 6) | (chan(g, 6) << 11) | (chan(b, 5) << 22))


@pytest.mark.parametrize("use_tonemap", [False, True])
def test_temporal_supersampling_accept_and_reject(ffi, oracle, use_tonemap):
    """temporalSupersampling.comp:86-110: an accepted history sample is blended 50:50, a rejected one (depth difference >= 1 m,
    2x2 luminance block contrast >= 0.5, reprojected position off screen) leaves the current sample."""
    w, h = 48, 32
    A, B = (0.5, 0.25, 0.125), (0.25, 0.75, 0.5)
    cur = np.full((h, w), pack_r11g11b10_exact(*A), np.uint32)
    last = np.full((h, w), pack_r11g11b10_exact(*B), np.uint32)
    zero_motion = np.zeros((h, w, 2), np.int16)
    near, far = 0.1, 300.0
    depth_of = lambda z: np.float32(1.0 - (near * far / z - far) / (near - far))  # inverse of linearDepth.inc:5-8
    d10, d12 = np.full((h, w), depth_of(10.0), np.float32), np.full((h, w), depth_of(12.0), np.float32)
    lum = np.full((h, w), 100, np.uint8)

    def tm(c):
        c = np.asarray(c, np.float64)
        return c / (1 + c @ np.array([0.21, 0.72, 0.07])) if use_tonemap else c

    def tm_inv(c):
        return c / (1 - c @ np.array([0.21, 0.72, 0.07])) if use_tonemap else c
    blended, kept = tm_inv(0.5 * tm(A) + 0.5 * tm(B)), np.asarray(A, np.float64)
    run = lambda **kw: decode_r11g11b10(passes.temporal_supersampling(ffi, oracle, **{**dict(current=cur, last=last, motion=zero_motion, depth_current=d10, depth_last=d10,
                                                                                           lum_current=lum, lum_last=lum, use_tonemap=use_tonemap), **kw}))
    tol = dict(rtol=2.0 ** -5, atol=0)  # R11G11B10 keeps 6 / 5 mantissa bits
    assert np.allclose(run(), blended, **tol)
    assert np.allclose(run(depth_last=d12), kept, **tol)                                   # depth test
    assert np.allclose(run(lum_current=np.full((h, w), 250, np.uint8), lum_last=np.full((h, w), 10, np.uint8)), kept, **tol)  # contrast test: 4 * (250 - 10) / 255 >= 0.5
    # the contrast is a difference of absolute values (:22-28), so a darker current block always passes
    assert np.allclose(run(lum_current=np.full((h, w), 10, np.uint8), lum_last=np.full((h, w), 250, np.uint8)), blended, **tol)
    off = zero_motion.copy()
    off[..., 0] = 32767  # motion of +1 screen width: every reprojected position is off screen
    assert np.allclose(run(motion=off), kept, **tol)
    # motion dilation picks the motion of the closest (largest, reverse z) depth in the 3x3 neighbourhood
    dil = zero_motion.copy()
    dil[10, 20, 0] = 32767
    dc = d10.copy()
    dc[10, 20] = depth_of(2.0)
    out = run(motion=dil, depth_current=dc, depth_last=dc)
    assert np.allclose(out[9:12, 19:22], kept, **tol) and np.allclose(out[0:8], blended, **tol)
