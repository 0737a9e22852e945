#!/bin/bash
# Final evidence of a round in one call: the bench line, the ncu launch list of the bench command, `ncu --set full` of the four big kernels
# (with source) and of every other pass kernel of one frame (raw CSV only: the report would exceed the 64 MiB gpurun_out limit).
# usage: gpu_final_profiles.sh <tag>
set -u
tag=$1; mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench: $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/${tag}_launches_bench.log 2>&1; echo "launch list: $?"
bash tools/gpu_ncu_big.sh ${tag}
REST='histogram|sky|Sky|preExpose|hiz|lightMatrix|depthDownscale|sdfFrustum|sdfTile|giTemporal|giUpscale|froxel|volum|bloom|applyBloom|tonemapping'
timeout 900 ncu --set full --clock-control none -k "regex:$REST" --launch-skip 128 -c 32 -f -o /tmp/${tag}_rest python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_rest_ncu.log 2>&1; echo "ncu rest: $?"
ncu -i /tmp/${tag}_rest.ncu-rep --page raw --csv > gpurun_out/${tag}_rest_raw.csv 2>/dev/null; ls -la gpurun_out/${tag}_rest_raw.csv
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("frames/s %.1f  ms %.3f  e2e %.1f  roofline %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: d["roofline"][k] for k in ("kernel", "frac", "kernel_ms")}))
print("pair", d["roofline_pair"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
