#!/bin/bash
# ncu --set full of selected kernels of one 4K frame (no graph replay). usage: gpu_ncu.sh <tag> <kernel-regex> [count]
set -u
tag=$1; K=$2; N=${3:-12}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip-before-match 0 -c $N -f -o gpurun_out/${tag} python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu: $?"
ls -la gpurun_out/${tag}.ncu-rep
