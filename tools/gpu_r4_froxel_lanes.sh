#!/bin/bash
# Last call of round 2 (about 100 s of box time): parity of the froxel chain with the power-of-two noise addressing (both the four kernels and the
# fused launch), then the bench with 64 / 32 / 16 z lanes per block of the fused launch, then the froxel parity cases for the two smaller blocks.
# usage: gpu_r4_froxel_lanes.sh <tag>
set -u
tag=$1; mkdir -p gpurun_out
t0=$(date +%s)
timeout 60 python -u -m pytest tests/test_zz_single_pass_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "froxel or moving_camera or (static_camera and 200)" > gpurun_out/${tag}_pytest64.log 2>&1
echo "pytest (64 lanes): rc $? at $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/${tag}_pytest64.log)"
for lanes in 64 32 16; do
    PLAIN_FROXEL_ZLANES=$lanes timeout 40 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/${tag}_bench_z${lanes}.json 2> gpurun_out/${tag}_bench_z${lanes}.err
    echo "bench $lanes lanes: rc $? at $(( $(date +%s) - t0 )) s"
done
python - <<PY
import json
for lanes in (64, 32, 16):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_z%d.json" % lanes))
        print(lanes, "lanes: frames/s %.1f  ms %.3f  froxel launch ms %.4f  trace %.4f" % (d["value"], d["ms_per_step"], d["passes_ms"]["Volumetric light integration"], d["passes_ms"]["Indirect diffuse SDF trace"]))
    except Exception as e:
        print(lanes, "unreadable:", e)
PY
for lanes in 32 16; do
    PLAIN_FROXEL_ZLANES=$lanes timeout 40 python -u -m pytest tests/test_zz_single_pass_gpu.py -m gpu -q -p no:cacheprovider -k froxel > gpurun_out/${tag}_pytest${lanes}.log 2>&1 &
done
wait
for lanes in 32 16; do echo "pytest ($lanes lanes) at $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/${tag}_pytest${lanes}.log)"; done
