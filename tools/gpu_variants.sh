#!/bin/bash
# A/B of scheduling variants of one kernel in one GPU call: `gpu_variants.sh <tag> <ENVVAR> v1 v2 ...` runs the bench once per value
set -u
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --no-cpu-baseline --steps 15 --warmup 4 > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_$v.json"))
p = d["passes_ms"]
print("$var=$v: frame %.3f ms | trace %.4f spatial %.4f taa %.4f shading %.4f bloomUp0 %.4f bloomDown1 %.4f" % (d["ms_per_step"], p["Indirect diffuse SDF trace"], p["Indirect diffuse spatial filter"], p["Temporal filtering"], p["Forward shading"], p["Bloom Upsample mip 0"], p["Bloom downsample mip 1"]))
PY
done
