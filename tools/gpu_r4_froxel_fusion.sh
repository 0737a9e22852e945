#!/bin/bash
# One B200 call for the froxel column fusion: the bench line with the fused launch (default), the whole GPU suite, the same bench with the four
# per-pass kernels (PLAIN_FROXEL_FUSION=0, A / B on the same box), `ncu --set full` of the new kernel. The suite writes its log as it goes (-v, unbuffered), so a call
# that runs into the box limit still leaves the results of the tests that ran. xdist workers share the one GPU (the tests wait on the CPU
# oracle most of the time); if xdist cannot start, the suite runs serially.
# usage: gpu_r4_froxel_fusion.sh <tag> [workers]
set -u
tag=$1; workers=${2:-3}; mkdir -p gpurun_out
t0=$(date +%s)
timeout 150 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench fused: rc $? at $(( $(date +%s) - t0 )) s"
timeout 290 python -u -m pytest tests -m gpu -v -p no:cacheprovider -n ${workers} --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
if [ $rc -ne 0 ] && ! grep -qE "PASSED|FAILED" gpurun_out/${tag}_pytest.log; then
    echo "xdist run did not start (rc $rc): serial run"
    timeout 290 python -u -m pytest tests -m gpu -v -p no:cacheprovider --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; rc=$?
fi
echo "pytest: rc $rc at $(( $(date +%s) - t0 )) s"
grep -E "FAILED|ERROR" gpurun_out/${tag}_pytest.log | head -40; echo "PASSED lines: $(grep -c PASSED gpurun_out/${tag}_pytest.log)"
tail -3 gpurun_out/${tag}_pytest.log
PLAIN_FROXEL_FUSION=0 timeout 100 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/${tag}_bench_unfused.json 2> gpurun_out/${tag}_bench_unfused.err; echo "bench unfused: rc $? at $(( $(date +%s) - t0 )) s"
python - <<PY
import json
for name in ("${tag}_bench", "${tag}_bench_unfused"):
    try:
        d = json.load(open("gpurun_out/%s.json" % name))
        p = d["passes_ms"]
        print(name, "frames/s %.1f  ms %.3f  e2e %.1f  launches/frame %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] // d["steps"]),
              "froxel chain ms", [p.get(k) for k in ("Froxel volume material", "Froxel light scattering", "Volumetric lighting reprojection", "Volumetric light integration")])
    except Exception as e:
        print(name, "unreadable:", e)
PY
timeout 90 ncu --set full --clock-control none --import-source on -k regex:froxelColumnKernel --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_froxel python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/${tag}_froxel_ncu.log 2>&1; echo "ncu froxelColumnKernel: rc $? at $(( $(date +%s) - t0 )) s"
ncu -i gpurun_out/${tag}_froxel.ncu-rep --page raw --csv > gpurun_out/${tag}_froxel_raw.csv 2>/dev/null; ls -la gpurun_out/${tag}_froxel_raw.csv
