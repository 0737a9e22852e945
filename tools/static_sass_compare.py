#!/usr/bin/env python
"""Static comparison of the exact and the fast build of the floating-point passes: SASS instruction counts and
the instruction classes that differ, per kernel, from `cuobjdump -sass` of plainrenderer_b200/_build/<unit>.o and <unit>_fast.o, plus
registers / spills from the ptxas logs. No GPU involved: these are counts of instructions in the binary, not of instructions executed.
usage: python tools/static_sass_compare.py > profiles/<name>.md"""
import collections
import re
import subprocess
from pathlib import Path

BUILD = Path(__file__).resolve().parents[1] / "plainrenderer_b200" / "_build"
UNITS = ["passes_gi", "passes_shading", "passes_post", "passes_volumetrics"]
CLASSES = [("MUFU", r"^MUFU"), ("FFMA", r"^FFMA"), ("FMUL/FADD", r"^(FMUL|FADD)"), ("FSETP/FSEL/FMNMX", r"^(FSETP|FSEL|FMNMX)"), ("int ALU", r"^(IADD3|IMAD|LOP3|SHF|LEA|ISETP|SEL|PRMT|IABS)"),
           ("cvt", r"^(F2I|I2F|F2F|FRND|F2FP|HADD2)"), ("ld/st", r"^(LDG|STG|LDS|STS|LDC|LD|ST|ATOM|RED)"), ("branch", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC)")]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    res, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("pb::", "")
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            res[cur]["total"] += 1
            for name, rx in CLASSES:
                if re.match(rx, op):
                    res[cur][name] += 1
    return res


def resources(log):
    out, cur = {}, None
    for line in Path(log).read_text().splitlines():
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("pb::", "")
        m = re.search(r"(\d+) bytes spill stores", line)
        if m and cur:
            out.setdefault(cur, {})["spill"] = int(m.group(1))
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            out.setdefault(cur, {})["regs"] = int(m.group(1))
    return out


print("# Exact vs fast build of the floating-point passes: static SASS (no GPU run)\n")
print("`python tools/static_sass_compare.py` - instruction counts in the sm_100a binaries (`cuobjdump -sass`), registers and spill bytes from `ptxas -v`.")
print("Static counts say how much code a contract needs, not how long a kernel runs; `libplain_b200_fast.so` has not been timed (DESIGN.md section 12).\n")
cols = ["total"] + [c for c, _ in CLASSES]
print("| kernel | build | regs | spill B | " + " | ".join(cols) + " |")
print("|---|---|---|---|" + "---|" * len(cols))
for unit in UNITS:
    a, b = kernels(BUILD / (unit + ".o")), kernels(BUILD / (unit + "_fast.o"))
    ra, rb = resources(BUILD / (unit + ".log")), resources(BUILD / (unit + "_fast.log"))
    for k in a:
        if a[k]["total"] < 300 or k not in b:
            continue
        for tag, c, r in (("exact", a[k], ra.get(k, {})), ("fast", b[k], rb.get(k, {}))):
            print("| `%s` | %s | %s | %s | " % (k[:60], tag, r.get("regs", ""), r.get("spill", "")) + " | ".join(str(c[x]) for x in cols) + " |")
