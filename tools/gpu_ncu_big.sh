#!/bin/bash
# ncu --set full (+ source correlation) of the four big kernels of one 4K frame of the current build. usage: gpu_ncu_big.sh <tag>
set -u
tag=$1; mkdir -p gpurun_out
BIG='sdfDiffuseTraceKernel|giSpatialFilterKernel|temporalFilterKernel|gbufferShadingKernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$BIG" --launch-skip 20 -c 5 -f -o gpurun_out/${tag}_big python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_big_ncu.log 2>&1; echo "ncu big: $?"
ls -la gpurun_out/${tag}_big.ncu-rep
