"""detmath.h (the pinned libm of the numeric contract) against float64 numpy: max error in ulp on the ranges the frame path uses."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def dm():
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libdetmath_c.so"
    src = ROOT / "tests" / "emul" / "detmath_c.cpp"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I", str(ROOT / "plainrenderer_b200" / "csrc"), str(src), "-o", str(lib)], check=True)
    return C.CDLL(str(lib))


def call1(dm, name, x):
    x = np.ascontiguousarray(x, np.float32)
    o = np.empty_like(x)
    getattr(dm, "dmw_" + name)(x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_int(x.size))
    return o


def call2(dm, name, x, y):
    x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
    o = np.empty_like(x)
    getattr(dm, "dmw_" + name)(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_int(x.size))
    return o


def ulp_error(got, want64):
    want32 = want64.astype(np.float32)
    spacing = np.spacing(np.abs(want32)).astype(np.float64)
    return np.max(np.abs(got.astype(np.float64) - want64) / spacing)


RNG = np.random.default_rng(7)


@pytest.mark.parametrize("name,fn,lo,hi,tol", [
    ("exp", np.exp, -80, 80, 2.0), ("exp2", np.exp2, -120, 120, 2.0), ("log", np.log, 1e-30, 1e30, 2.0), ("log2", np.log2, 1e-30, 1e30, 2.5),
    ("sin", np.sin, -100, 100, 2.5), ("cos", np.cos, -100, 100, 2.5), ("asin", np.arcsin, -1, 1, 3.0), ("acos", np.arccos, -1, 1, 3.0), ("atan", np.arctan, -1e4, 1e4, 2.5)])
def test_unary(dm, name, fn, lo, hi, tol):
    if name in ("log", "log2"):
        x = np.exp(RNG.uniform(np.log(lo), np.log(hi), 200000)).astype(np.float32)
    else:
        x = RNG.uniform(lo, hi, 200000).astype(np.float32)
    got = call1(dm, name, x)
    want = fn(x.astype(np.float64))
    if name in ("sin", "cos"):  # absolute error near zeros of the function
        assert np.max(np.abs(got - want)) < 2e-7 * 4
    else:
        assert ulp_error(got, want) <= tol


def test_pow(dm):
    x = np.exp(RNG.uniform(-10, 10, 200000)).astype(np.float32)
    y = RNG.uniform(-8, 8, 200000).astype(np.float32)
    got = call2(dm, "pow", x, y)
    want = np.power(x.astype(np.float64), y.astype(np.float64))
    ok = np.isfinite(want) & (want > 1e-30) & (want < 1e30)
    rel = np.abs(got[ok] - want[ok]) / want[ok]
    assert rel.max() < 2e-5  # exp2(y*log2(x)) amplifies the log2 rounding by |y*log2 x|
    # pinned edge cases of the numeric contract
    assert call2(dm, "pow", [-0.5, 0.0, 0.0, 2.0], [5.0, 2.0, 0.0, 0.0]).tolist() == [0.0, 0.0, 1.0, 1.0]


def test_special_values(dm):
    assert np.isnan(call1(dm, "log", [-1.0]))[0]
    assert call1(dm, "log", [0.0])[0] == -np.inf
    assert call1(dm, "exp", [-200.0, 200.0]).tolist() == [0.0, np.inf]
    assert call1(dm, "acos", [1.0000001, -1.0000001]).tolist() == pytest.approx([0.0, np.pi], abs=1e-6)
    assert call2(dm, "atan2", [0.0, 1.0, -1.0], [0.0, 0.0, 0.0]).tolist() == pytest.approx([0.0, np.pi / 2, -np.pi / 2], abs=1e-6)
