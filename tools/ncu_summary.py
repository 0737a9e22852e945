#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv > raw.csv; ncu_summary.py raw.csv [out.json]`.
Columns: duration, DRAM bytes (read + write), warp instructions, issue-slot utilisation, lanes per instruction, registers,
local (spill) bytes, achieved occupancy, L1/L2 hit rates, top stall reasons."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def f(r, name, default=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def scale(name, v):
    u = units[col[name]] if name in col else ""
    k = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}
    return v * k.get(u, 1) if v is not None else None


out = []
for r in data:
    name = r[col["Kernel Name"]]
    stalls = {h.split("_pipe_stalled_")[-1] if False else h: f(r, h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    stalls = sorted(((v, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k, v in stalls.items() if v), reverse=True)[:4]
    e = {
        "kernel": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
        "ms": scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum")),
        "dram_bytes": (scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum", 0)) or 0) + (scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum", 0)) or 0),
        "warp_inst": f(r, "smsp__inst_executed.sum"),
        "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "regs": f(r, "launch__registers_per_thread"),
        "local_load_bytes": f(r, "smsp__inst_executed_op_local_ld.sum"),
        "local_store_inst": f(r, "smsp__inst_executed_op_local_st.sum"),
        "occupancy_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "theoretical_occ_pct": f(r, "sm__maximum_warps_per_active_cycle_pct"),
        "l1_hit_pct": f(r, "l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
        "smem_per_block": f(r, "launch__shared_mem_per_block_static"),
        "stalls_per_issue": [[k, round(v, 2)] for v, k in stalls],
    }
    out.append(e)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
print("| kernel | grid | ms | DRAM MB | warp inst (M) | issue % | lanes/inst | regs | local ld/st inst | occ % (theo) | L1 / L2 hit % | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for e in out:
    print("| `%s` | %s | %.4f | %.1f | %.1f | %.1f | %.1f | %d | %s / %s | %.0f (%.0f) | %.0f / %.0f | %s |" % (
        e["kernel"][:60], e["grid"], e["ms"] or 0, (e["dram_bytes"] or 0) / 1e6, (e["warp_inst"] or 0) / 1e6, e["issue_active_pct"] or 0, e["lanes_per_inst"] or 0, e["regs"] or 0,
        int(e["local_load_bytes"] or 0), int(e["local_store_inst"] or 0), e["occupancy_pct"] or 0, e["theoretical_occ_pct"] or 0, e["l1_hit_pct"] or 0, e["l2_hit_pct"] or 0,
        ", ".join("%s %.2f" % (k, v) for k, v in e["stalls_per_issue"])))
