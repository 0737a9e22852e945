#!/usr/bin/env python
"""Warp instructions and stall samples per CUDA source line of one kernel, from
`ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1`.
usage: line_mix.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= iE or not r[0].strip().isdigit():
        continue
    try:
        out.append((int(r[iE]), int(r[iS]), int(r[iT]), cur_file, int(r[0]), r[1].strip()))
    except ValueError:
        pass
tot = sum(o[0] for o in out) or 1
ts = sum(o[1] for o in out) or 1
print("total warp instructions attributed to lines:", tot)
for e, s, t, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.2f%% inst %5.2f%% samples  lanes %4.1f  %s:%d  %s" % (100 * e / tot, 100 * s / ts, t / max(e, 1), f, ln, src[:90]))
