"""bloomDownsample.comp / bloomUpsample.comp / applyBloom.comp (S12) of the oracle against an independent float64 numpy restatement
written from the GLSL: the 13-tap Call-of-Duty downsample, the 9-tap tent + 4-tap box upsample with the push-constant blur radius, the
lowest-mip specialisation, mix(scene, bloom, strength). Every level is recomputed from the ORACLE's previous level, so each comparison
isolates one dispatch; a result must be the R11G11B10 value nearest to the float64 one (round to nearest: half a mantissa step)."""
import numpy as np
import pytest

import passes
from conftest import decode_r11g11b10, random_r11g11b10
from test_gi_temporal_upscale_numpy import bilinear


def level(packed, w, h):
    return decode_r11g11b10(np.asarray(packed, np.uint32).reshape(h, w))


def taps(src, tw, th, offsets, step):
    """sum of weight * textureLod(src, uv + step * offset) over the target grid tw x th; step in uv units (x, y)"""
    ys, xs = np.mgrid[0:th, 0:tw]
    u, v = (xs + 0.5) / tw, (ys + 0.5) / th
    acc = np.zeros((th, tw, 3))
    for (ox, oy), wt in offsets:
        acc += bilinear(src, u + step[0] * ox, v + step[1] * oy) * wt
    return acc


DOWN = [((0, 0), 0.125)] + [((x, y), 0.125) for x in (0.5, -0.5) for y in (0.5, -0.5)] + [((1.5, 0), 0.0625), ((-1.5, 0), 0.0625), ((0, 1.5), 0.0625), ((0, -1.5), 0.0625)] \
    + [((x, y), 0.03125) for x in (1.5, -1.5) for y in (1.5, -1.5)]                                                     # bloomDownsample.comp:28-48
TENT = [((0, 0), 0.25)] + [((1, 0), 0.125), ((-1, 0), 0.125), ((0, 1), 0.125), ((0, -1), 0.125)] + [((x, y), 0.0625) for x in (1, -1) for y in (1, -1)]  # bloomUpsample.comp:37-47
BOX = [((x, y), 0.25) for x in (0.5, -0.5) for y in (0.5, -0.5)]                                                        # bloomUpsample.comp:52-55


def assert_nearest_code(got, want, what):
    """got: decoded R11G11B10 texels; want: float64. Round-to-nearest: within half a step (2^-7 of the value for the 6-bit mantissas,
    2^-6 for blue), plus a sliver for the oracle's binary32 accumulation."""
    step = np.array([2.0 ** -7, 2.0 ** -7, 2.0 ** -6]) * 1.02
    floor = np.array([2.0 ** -14, 2.0 ** -14, 2.0 ** -14])  # below the smallest normal the step is absolute
    err = np.abs(got - want) / np.maximum(np.abs(want), floor)
    assert (err <= step).all(), "%s: %d texels off, worst %.4f steps" % (what, int((err > step).any(-1).sum()), float((err / step).max()))


@pytest.mark.parametrize("w,h,radius,strength", [(64, 48, 1.5, 0.05), (100, 75, 1.0, 0.3), (33, 17, 2.5, 1.0), (96, 64, 1.5, 0.0)])
def test_bloom_chain_matches_numpy(ffi, oracle, w, h, radius, strength):
    rng = np.random.default_rng(w + h)
    # HDR scene: mid-grey noise with a few very bright texels (what bloom is for); exponents 9..20 keep everything normalised
    def chan(mbits, lo, hi):
        return (rng.integers(lo, hi, (h, w), dtype=np.uint32) << mbits) | rng.integers(0, 1 << mbits, (h, w), dtype=np.uint32)
    packed = chan(6, 11, 16) | (chan(6, 11, 16) << 11) | (chan(5, 11, 16) << 22)
    hot = rng.random((h, w)) < 0.01
    packed[hot] = (22 << 6) | ((21 << 6) << 11) | ((20 << 5) << 22)
    mips = 6
    down, up, result = passes.bloom(ffi, oracle, packed, strength=strength, radius=radius, mips=mips)
    res = lambda m: (max(w >> m, 1), max(h >> m, 1))
    scene = decode_r11g11b10(packed)
    # downsample chain: mip m from mip m - 1 (mip 0 = the scene)
    src = scene
    for m in range(1, mips):
        tw, th = res(m)
        sh, sw = src.shape[:2]
        want = taps(src, tw, th, DOWN, (1.0 / sw, 1.0 / sh))
        got = level(down[m - 1], tw, th)
        assert_nearest_code(got, want, "downsample mip %d" % m)
        src = got
    # upsample chain: target mip t from down[t + 1] (tent, step = radius / source size) + up[t + 1] (box, step = 1 / source size)
    for i in range(mips - 1):
        t = mips - 2 - i
        tw, th = res(t)
        sw, sh = res(t + 1)
        source = level(down[t], sw, sh)                     # down[] holds mips 1..5: index t = mip t + 1
        want = taps(source, tw, th, TENT, (radius / sw, radius / sh))
        if i > 0:                                           # isLowestMip = false
            want += taps(level(up[t + 1], sw, sh), tw, th, BOX, (1.0 / sw, 1.0 / sh))
        assert_nearest_code(level(up[t], tw, th), want, "upsample mip %d" % t)
    # apply: mix(scene, bloom, strength) in place
    ys, xs = np.mgrid[0:h, 0:w]
    bloom = bilinear(level(up[0], w, h), (xs + 0.5) / w, (ys + 0.5) / h)
    assert_nearest_code(decode_r11g11b10(result), scene * (1 - strength) + bloom * strength, "apply")
    if strength == 0.0:
        assert np.array_equal(result, packed)
    assert level(up[0], w, h).max() > scene[~hot].max()    # the hot texels did bleed
