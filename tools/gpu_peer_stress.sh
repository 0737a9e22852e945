#!/bin/bash
# repeated peer-exchange parity runs with host-side skew on either rank. usage: gpu_peer_stress.sh <tag> <N>
tag=$1; N=$2; mkdir -p gpurun_out; port=29600
for extra in "" "" "--skew=1:20" "--skew=0:20" "--skew=1:3"; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/check_sharded_nccl.py 512 512 10 $extra 2>&1 | grep "differs\|SHARDED\|rror" | head -6
done
