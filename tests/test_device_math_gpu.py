"""Exhaustive on-device checks (include/plain_b200.h plain_device_selftest) of the instruction-lean sequences of round 2 against the
functions of the numeric contract they replace: the single-test and untested reciprocal paths against __frcp_rn over every binary32
value of their domain, the lean R11G11B10 codecs against the contract's over all 2^32 values / all codes ON THE DEVICE (the host side
is tests/test_codecs.py), one-conversion floor, FMNMX against the pinned min / max where no operand is -0."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

NAMES = ["rcpf_nz vs __frcp_rn", "rcpf_normal vs __frcp_rn", "lean R11G11B10 encoder", "lean R11G11B10 decoder", "floor2i + int->float vs floor_ + f2i",
         "FMNMX vs pinned min/max (no -0)", "sqrtf_normal / sqrt2_normal vs sqrtf (2^-60..2^60)", "divf_normal / div2_normal / rcp2_normal vs IEEE division / __frcp_rn (2^-60..2^60)"]


def test_lean_device_sequences_equal_the_contract_functions(ffi, cuda):
    be = ffi.Backend(cuda, 0, 64, 64)
    out = (C.c_uint64 * 8)()
    be._check(cuda.b["device_selftest"](be.ctx, out), "device_selftest")
    counts = list(out)
    be.close()
    for name, n in zip(NAMES, counts):
        assert n == 0, "%s: %d mismatching bit patterns" % (name, n)
