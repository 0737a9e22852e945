"""The tolerance of the "fast" numeric contract, shared by the GPU test of libplain_b200_fast.so and by
the CPU test of its error model (liboracle_sfu.so): both compare a frame sequence with the exact oracle's through these bounds."""
import numpy as np

from conftest import decode_r11g11b10

GI_IMAGES = ["giY1", "giC1", "giFullY", "giFullC"]                            # RGBA16F / RG16F: denoised + upscaled sphere-trace results
HALF_IMAGES = ["froxelIntegration"]                                           # RGBA16F, smooth
PACKED_IMAGES = ["color0", "color1", "taaHist0", "taaHist1", "post0"]         # R11G11B10
EXACT_IMAGES = ["hiz", "depthHalf"]                                           # min / max and point sampling of uploaded depth


def rel_error(a, b, floor):
    """|a - b| relative to max(|b|, floor): floor = the magnitude below which a difference cannot reach the 8-bit frame."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    assert (np.isfinite(a) == np.isfinite(b)).all(), "inf / NaN pattern differs"
    ok = np.isfinite(b)
    return np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), floor)


def check(name, e, p999_max, mean_max, log):
    p999, mean, worst = float(np.quantile(e, 0.999)), float(e.mean()), float(e.max())
    log.append("%-20s mean %.2e  p99.9 %.2e  max %.2e" % (name, mean, p999, worst))
    return [] if (p999 <= p999_max and mean <= mean_max) else ["%s: mean %.2e (<= %.0e), p99.9 %.2e (<= %.0e)" % (name, mean, mean_max, p999, p999_max)]


def check_outliers(name, e, over, frac_max, mean_max, log):
    frac, mean = float((e > over).mean()), float(e.mean())
    log.append("%-20s mean %.2e  above %.0e: %.3f %% of the texels, max %.2e" % (name, mean, over, 100 * frac, float(e.max())))
    return [] if (frac <= frac_max and mean <= mean_max) else ["%s: mean %.2e (<= %.0e), %.3f %% above %.0e (<= %.1f %%)" % (name, mean, mean_max, 100 * frac, over, 100 * frac_max)]


def compare_snapshots(sa, sb):
    """sa: the approximate library, sb: the exact oracle (conftest.Sequence.snapshot()). Returns (failures, log lines)."""
    bad, log = [], []
    for k in sa:
        if k.split("/")[0] in EXACT_IMAGES and not np.array_equal(sa[k], sb[k]):
            bad.append("%s must be bit-exact" % k)
    d = np.abs(sa["output/0"].astype(np.int32) - sb["output/0"].astype(np.int32))
    frac_le1 = float((d <= 1).mean())
    log.append("8-bit frame: max |diff| %d / 255, <= 1 LSB on %.4f %% of the bytes, mean %.4f" % (d.max(), 100 * frac_le1, d.mean()))
    if not (frac_le1 >= 0.999 and d.max() <= 4):
        bad.append("8-bit frame: <= 1 LSB on %.4f %% (>= 99.9), max %d (<= 4)" % (100 * frac_le1, d.max()))
    for name in PACKED_IMAGES:
        ea, eb = decode_r11g11b10(sa[name + "/0"].view(np.uint32)), decode_r11g11b10(sb[name + "/0"].view(np.uint32))
        # one step of the 6 / 5 bit mantissas is 1.6e-2 / 3.1e-2: most texels identical, a few one or two steps apart
        # (bounds with margin over the error model on a 100-instance scene with a moving camera: p99.9 4.4e-2, mean 5e-4)
        bad += check(name, rel_error(ea, eb, 1e-3 * float(np.median(eb[np.isfinite(eb)])) + 1e-12), 6e-2, 2e-3, log)
    for name in HALF_IMAGES + GI_IMAGES:
        ha, hb = sa[name + "/0"].view(np.float16), sb[name + "/0"].view(np.float16)
        scale = float(np.abs(hb[np.isfinite(hb)].astype(np.float64)).mean()) + 1e-12
        e = rel_error(ha, hb, 1e-2 * scale)
        if name in GI_IMAGES:
            # the sphere trace is a discrete process: an error of an ulp flips a few rays between hit and miss (or between two
            # instances), which the denoiser spreads over their neighbourhood - the median error is 0, a fraction of a percent of
            # the texels are simply different. Bounded as a fraction of outliers + the mean, not as a quantile
            bad += check_outliers(name, e, 2e-2, 0.05, 1e-2, log)   # error model, 100 instances, moving: 3.2 % outliers, mean 5.5e-3
        else:
            bad += check(name, e, 2e-2, 2e-3, log)
    # exposure follows the histogram of the previous frame: same bins up to the texels that moved across a bin edge
    hist_a, hist_b = sa["buf:histogram"].view(np.uint32).astype(np.int64), sb["buf:histogram"].view(np.uint32).astype(np.int64)
    moved = int(np.abs(hist_a - hist_b).sum())
    log.append("histogram: %d of %d counts moved" % (moved, int(hist_b.sum())))
    if moved > 0.01 * hist_b.sum():
        bad.append("histogram: %d counts moved" % moved)
    light_a, light_b = sa["buf:light"].view(np.float32), sb["buf:light"].view(np.float32)
    if not np.allclose(light_a, light_b, rtol=2e-3, atol=1e-12):
        bad.append("light buffer %s vs %s" % (light_a, light_b))
    return bad, log


def run_sequence(ffi, approx_api, oracle_api, moving, w=256, h=144, frames=6, instances=16, **settings):
    """Returns the comparison after the LAST frame: every history (TAA, GI, froxels, exposure) has been fed back frames - 1 times."""
    from conftest import Sequence
    a, b = Sequence(ffi, approx_api, w, h, instances, **settings), Sequence(ffi, oracle_api, w, h, instances, **settings)
    try:
        for _ in range(frames):
            inputs = a.step(moving=moving)
            b.step(moving=moving, inputs=inputs)
        return compare_snapshots(a.snapshot(), b.snapshot())
    finally:
        a.close()
        b.close()
