#!/bin/bash
# ncu --set full of ONE whole 4K frame of the current build (no graph replay), in two reports: the four big kernels with source
# correlation, every other pass kernel without. usage: gpu_ncu_frame.sh <tag>
set -u
tag=$1
mkdir -p gpurun_out
BIG='sdfDiffuseTraceKernel|giSpatialFilterKernel|temporalFilterKernel|gbufferShadingKernel'
# 5 big launches per frame (two spatial filters); skip the first 4 frames
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$BIG" --launch-skip 20 -c 5 -f -o gpurun_out/${tag}_big python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_big_ncu.log 2>&1; echo "ncu big: $?"
REST='histogram|sky|Sky|preExpose|hiz|lightMatrix|depthDownscale|sdfFrustum|sdfTile|giTemporal|giUpscale|froxel|volum|bloom|applyBloom|tonemapping'
timeout 900 ncu --set full --clock-control none -k "regex:$REST" --launch-skip 132 -c 33 -f -o gpurun_out/${tag}_rest python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_rest_ncu.log 2>&1; echo "ncu rest: $?"
ls -la gpurun_out/${tag}_big.ncu-rep gpurun_out/${tag}_rest.ncu-rep
