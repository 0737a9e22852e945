#!/usr/bin/env python
"""Instruction mix of one kernel from an `ncu --page source --csv --print-source sass` export (first launch in the file).
usage: sass_mix.py file.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
op, samp, tot, static, kernels = collections.Counter(), collections.Counter(), 0, 0, 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        continue
    if r and r[0] == "Address":
        hdr = r
        iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or kernels > 1 or len(r) <= iE:
        continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    m = s.split()[0].rstrip(";") if s else ""
    base = m.split(".")[0]
    if base in ("MUFU", "F2I", "I2F", "FRND", "F2F", "LDG", "LDS", "STS", "LDC"):
        base = ".".join(m.split(".")[:2])
    e = int(r[iE])
    op[base] += e
    tot += e
    static += 1
    samp[base] += int(r[iSm])
print("total warp instructions", tot, "static", static)
ss = max(sum(samp.values()), 1)
for k, v in op.most_common(top):
    print("%-14s %6.2f%%  samples %5.1f%%" % (k, 100 * v / tot, 100 * samp[k] / ss))
