#!/bin/bash
# Quick edit-measure loop: frame parity + single-pass cases, then the bench line. usage: gpu_quick.sh <tag> [extra pytest args]
set -u
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_zz_single_pass_gpu.py -q -x -m gpu --durations=8 "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest: $?"
tail -14 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench: $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("frames/s %.1f  ms %.3f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    print({k: v for k, v in d["passes_ms"].items() if v > 0.04})
except Exception as e:
    print("no bench line:", e)
PY
