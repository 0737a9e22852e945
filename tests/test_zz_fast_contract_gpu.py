"""The "fast" numeric contract (libplain_b200_fast.so): the floating-point passes of the frame path with
the SFU approximations (ex2 / lg2 / sin / cos / rcp / rsqrt / sqrt .approx) and compiler contraction instead of the pinned
binary32 sequences. Parity is then a TOLERANCE against the oracle - the north star's "within 1e-3 relative L-inf on the
final tonemapped frame" read on the frame's own scale (a unit of 1/255 is 3.9e-3: the dithered 8-bit frame cannot be closer
than 0 or 1 LSB), plus bounds on the HDR intermediates - and the integer passes stay bit-exact on the same inputs.

First seen green on a B200 in round 2 (profiles/r2a_staged_tests.md). The default contract of bench.py and of every other test is
"exact"; the fast library is an experiment reported beside the headline, never instead of it."""
import numpy as np
import pytest

import passes
import tolerance
from conftest import random_r11g11b10

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("moving", [False, True])
def test_fast_frame_sequence_within_tolerance(ffi, cuda_fast, oracle, moving):
    bad, log = tolerance.run_sequence(ffi, cuda_fast, oracle, moving)
    print("\n".join("fast-contract " + l for l in log))
    assert not bad, "fast contract outside its tolerance: " + "; ".join(bad)


@pytest.mark.parametrize("settings", [dict(indirect_lighting_tech=1), dict(diffuse_brdf=3, direct_multiscatter=3), dict(half_res_trace=0), dict(sun_shadow_cascade_count=4),
                                      dict(taa_history_sampling_tech=1), dict(taa_use_separate_supersampling=1), dict(strict_influence_radius_cutoff=1)])
def test_fast_setting_variants_within_tolerance(ffi, cuda_fast, oracle, settings):
    """the non-default configurations (all inside the tolerance under the CPU error model)"""
    bad, log = tolerance.run_sequence(ffi, cuda_fast, oracle, True, w=160, h=90, frames=4, **settings)
    print("\n".join("fast-contract %s " % settings + l for l in log[:1]))
    assert not bad, "fast contract outside its tolerance with %s: %s" % (settings, "; ".join(bad))


def test_fast_integer_passes_stay_bit_exact(ffi, cuda_fast, oracle):
    rng = np.random.default_rng(11)
    depth = rng.random((70, 130), dtype=np.float32)
    depth[rng.random((70, 130)) < 0.2] = 0.0
    for x, y in zip(passes.hiz(ffi, cuda_fast, depth), passes.hiz(ffi, oracle, depth)):
        assert np.array_equal(x, y)
    packed = random_r11g11b10(rng, 96 * 70).reshape(70, 96)
    (ta, ha), (tb, hb) = passes.histogram(ffi, cuda_fast, packed, 0.37), passes.histogram(ffi, oracle, packed, 0.37)
    assert np.array_equal(ta, tb) and np.array_equal(ha, hb)


def test_fast_tonemap_within_one_lsb(ffi, cuda_fast, oracle):
    rng = np.random.default_rng(12)
    packed = random_r11g11b10(rng, 130 * 67).reshape(67, 130)
    a, b = passes.tonemap(ffi, cuda_fast, packed).astype(np.int32), passes.tonemap(ffi, oracle, packed).astype(np.int32)
    assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.05
