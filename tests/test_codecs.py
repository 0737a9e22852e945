"""The instruction-lean R11G11B10 codecs the "fast" contract's device code uses (image_view.h encodeSmallFloatFast / decodeSmallFloatFast,
DESIGN.md section 12) against the contract's codecs: all 2 x 2048 codes for the decoder, all 2^32 binary32 values x both mantissa widths
for the encoder - they are the same functions, so the fast library stores the nearest code exactly like the exact one."""
import os
import subprocess

from conftest import ROOT


def test_fast_codecs_equal_the_contract_codecs_exhaustively():
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "codec_check"
    src = ROOT / "tests" / "emul" / "codec_check.cpp"
    hdr = ROOT / "plainrenderer_b200" / "csrc" / "image_view.h"
    if not exe.exists() or exe.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-ffp-contract=off", "-I%s" % (ROOT / "plainrenderer_b200" / "csrc"), "-I%s" % (ROOT / "include"),
                        str(src), "-o", str(exe), "-lpthread"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "decode mismatches: 0" in r.stdout and "x 2 formats: 0" in r.stdout
