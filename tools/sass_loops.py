#!/usr/bin/env python
"""Static view of a kernel's loops without a GPU: `cuobjdump -sass -fun <mangled-substring> lib.so | sass_loops.py`
lists every backward branch (loop) with the number of instructions in its body and the body's opcode histogram."""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            body = ins[addr_index[tgt]:i + 1]
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
            print("loop 0x%04x..0x%04x: %d instructions  %s" % (tgt, a, len(body), ", ".join("%s %d" % kv for kv in ops.most_common(14))))
