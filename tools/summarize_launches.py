"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table of the LAST frame
(the last `--frame-launches` launches): usage  python tools/summarize_launches.py gpurun_out/launches.csv 37 > profiles/x.md"""
import collections
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
last = rows[-n:]
tot = sum(v for _, v in last)
agg = collections.OrderedDict()
cnt = collections.Counter()
for k, v in last:
    agg[k] = agg.get(k, 0) + v
    cnt[k] += 1
print("| kernel | launches | time (us) | share of frame |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k.split("(")[0].replace("void ", ""), cnt[k], v / 1000, 100 * v / tot))
print("| **frame total (%d launches, serialised, cold cache)** | | **%.1f** | |" % (n, tot / 1000))
