#!/bin/bash
# One call on N GPUs: peer-exchange parity, the headline bench, and BASELINE configs[4] (7680x4320, 256 instances). usage: gpu_n8.sh <tag> <N>
set -u
tag=$1; N=$2; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $T --master-port 29711 tools/check_sharded_nccl.py 1024 1024 6 > gpurun_out/${tag}_parity_n$N.log 2>&1; echo "parity: $?"; grep "SHARDED_\|differs" gpurun_out/${tag}_parity_n$N.log | head -4
timeout 300 $T --master-port 29712 bench.py --gpus $N --steps 100 --warmup 8 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench: $?"
timeout 400 $T --master-port 29713 bench.py --gpus $N --steps 40 --warmup 6 --workload c5 > gpurun_out/${tag}_bench_c5_n$N.json 2> gpurun_out/${tag}_bench_c5_n$N.err; echo "bench c5: $?"
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_n$N.json", "gpurun_out/${tag}_bench_c5_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f frames/s  %.3f ms  e2e %.1f frames/s  clock samples %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["samples"]))
        print({k: v for k, v in d["passes_ms"].items() if v > 0.03})
    except Exception as e:
        print(f, "no bench line:", e)
PY
