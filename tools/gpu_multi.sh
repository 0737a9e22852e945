#!/bin/bash
# Row-sharded frame on N GPUs of one box: bit-exact parity against the unsharded frame over peer exchange, then the bench with the
# next-frame exchanges deferred behind the frame (default) and synchronous (PLAIN_PEER_DEFERRED=0). usage: gpu_multi.sh <tag> <N> [steps]
set -u
tag=$1; N=$2; K=${3:-60}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_nccl.py 512 512 6 > gpurun_out/${tag}_parity_n$N.log 2>&1; echo "parity: $?"; grep "SHARDED_\|DIFFERS" gpurun_out/${tag}_parity_n$N.log | head -5
for d in 1 0; do
  PLAIN_PEER_DEFERRED=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$d bench.py --gpus $N --steps $K --warmup 8 > gpurun_out/${tag}_bench_n${N}_deferred$d.json 2> gpurun_out/${tag}_bench_n${N}_deferred$d.err; echo "bench deferred=$d: $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n${N}_deferred$d.json").read().strip().splitlines()[-1])
    print("N=$N deferred=$d: %.1f frames/s  %.3f ms  e2e %.1f frames/s  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["samples"]))
except Exception as e:
    print("no bench line:", e)
PY
done
