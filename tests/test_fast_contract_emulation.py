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


def test_approximate_reciprocals_are_rejected_for_a_reason(ffi, oracle_sfu, oracle):
    """Why the fast contract keeps 1 / x correctly rounded: sdfDiffuseTrace.comp:120 samples depth and normals with NEAREST filtering at
    uv = iUV / size - exactly on texel borders. With rcp.approx (one ulp) in that division a large part of the traced pixels read the
    neighbouring texel; with the adopted model (same frame, same inputs) a fraction of a percent of the rays differ."""
    lib = ROOT / "oracle" / "_build" / "liboracle_sfu_rcp.so"
    subprocess.run(["make", "-C", str(ROOT / "oracle"), str(lib)], check=True, capture_output=True)
    rejected = ffi.Api(str(lib), "oracle_", "oracle_frontend_")
    import passes
    from test_sdf_diffuse_trace_numpy import wall_inputs
    rng = np.random.default_rng(3)
    w, h = 96, 56                                                # 1 / 96 and 1 / 56 are not representable: x * (1 / w) * w straddles x
    _, _, noise, sky_p = wall_inputs(rng, 64, 40)
    near, far = 0.1, 300.0
    depth = (1 - (near * far / rng.uniform(6.0, 14.0, (h, w)) - far) / (near - far)).astype(np.float32)   # every texel its own depth
    normal = np.zeros((h, w, 4), np.uint8)
    normal[..., :3] = np.round((np.array([0.0, 0.0, 1.0]) * 0.5 + 0.5) * 255)
    run = lambda api: passes.sdf_diffuse_trace(ffi, api, depth, normal, noise, sky_p, [], [], np.zeros((8, 8), np.uint16), np.eye(4).ravel(), [1, 1, 1, 1, 1])[0].astype(np.float64)
    ref, adopted, rej = run(oracle), run(oracle_sfu), run(rejected)
    scale = 1e-2 * np.abs(ref).mean()
    frac = lambda x: float((np.abs(x - ref) / np.maximum(np.abs(ref), scale) > 2e-2).any(-1).mean())
    print("traced texels further than 2e-2 from the exact oracle: adopted model %.4f, with rcp.approx %.4f" % (frac(adopted), frac(rej)))
    assert frac(adopted) < 0.02 and frac(rej) > 0.10
