"""Committed golden vectors (tests/golden/frame_sequence.json, generated from the oracle by tests/golden/make_golden.py)."""
import importlib.util
import json
from pathlib import Path

import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


def _generator():
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _check(got, want, who):
    for f, (a, b) in enumerate(zip(got["frames"], want["frames"])):
        assert a["histogram"] == b["histogram"], "%s: frame %d luminance histogram" % (who, f)
        assert a["light"] == b["light"], "%s: frame %d LightBuffer" % (who, f)
        assert a["output_patch"] == b["output_patch"], "%s: frame %d tonemapped patch" % (who, f)
        bad = [k for k in b["sha256"] if a["sha256"][k] != b["sha256"][k]]
        assert not bad, "%s: frame %d resources differ from the golden vectors: %s" % (who, f, bad)


def test_oracle_reproduces_golden(oracle):
    want = json.loads((GOLDEN / "frame_sequence.json").read_text())
    _check(_generator().run(oracle), want, "oracle")


@pytest.mark.gpu
def test_cuda_reproduces_golden(cuda):
    want = json.loads((GOLDEN / "frame_sequence.json").read_text())
    _check(_generator().run(cuda), want, "CUDA")


def _check_n4(got, want, who):
    assert [g["settings"] for g in got] == [w["settings"] for w in want]
    for g, w in zip(got, want):
        for f, (a, b) in enumerate(zip(g["frames"], w["frames"])):
            bad = [k for k in b if a[k] != b[k]]
            assert not bad, "%s: %s frame %d differs from the golden vectors: %s" % (who, w["settings"], f, bad)


def test_oracle_reproduces_n4_golden(oracle):
    """temporal supersampling + SDF debug visualiser variants, tests/golden/n4_variants.json"""
    want = json.loads((GOLDEN / "n4_variants.json").read_text())
    _check_n4(_generator().run_n4(oracle), want, "oracle")


@pytest.mark.gpu
def test_cuda_reproduces_n4_golden(cuda):
    want = json.loads((GOLDEN / "n4_variants.json").read_text())
    _check_n4(_generator().run_n4(cuda), want, "CUDA")
