"""CPU evidence for the tolerance of the "fast" contract. liboracle_sfu.so is the oracle with its
floating-point passes compiled under an error model of the fast build - transcendentals displaced by the documented error of
the SFU instruction that replaces them (oracle/sfu_emulation.h), contraction left to the compiler - and must stay inside the
same bounds (tests/tolerance.py) the GPU test holds libplain_b200_fast.so to: errors of that size, fed back through the TAA /
GI / froxel / exposure histories over a frame sequence, do not grow. A model of error magnitudes, not of the hardware's bits."""
import subprocess

import numpy as np
import pytest

import passes
import tolerance
from conftest import ROOT, random_r11g11b10


@pytest.fixture(scope="module")
def oracle_sfu(ffi, oracle):
    lib = ROOT / "oracle" / "_build" / "liboracle_sfu.so"
    if not lib.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return ffi.Api(str(lib), "oracle_", "oracle_frontend_")


def test_error_model_stays_inside_the_tolerance(ffi, oracle_sfu, oracle):
    bad, log = tolerance.run_sequence(ffi, oracle_sfu, oracle, moving=True, w=192, h=108, frames=5)
    print("\n".join(log))
    assert not bad, "; ".join(bad)
    # ... and the model really perturbs: the HDR colour is not the exact oracle's
    assert any(l.startswith("color") and "mean 0.00e+00" not in l for l in log)


def test_error_model_tonemap_within_one_lsb(ffi, oracle_sfu, oracle):
    rng = np.random.default_rng(12)
    packed = random_r11g11b10(rng, 130 * 67).reshape(67, 130)
    a, b = passes.tonemap(ffi, oracle_sfu, packed).astype(np.int32), passes.tonemap(ffi, oracle, packed).astype(np.int32)
    assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.05
