#!/bin/bash
# Final check of round 2 (about 40 s of box time) with the fused froxel launch at its default of 16 z lanes: the froxel cases, two frame
# sequences (every image against the oracle, fused and unfused), the 4-rank sharded frame with pass fusion, then the bench line.
set -u
tag=$1; mkdir -p gpurun_out
t0=$(date +%s)
timeout 45 python -u -m pytest tests/test_zz_single_pass_gpu.py tests/test_parity_gpu.py tests/test_sharding_gpu.py -m gpu -q -p no:cacheprovider -k "froxel or moving_camera or (static_camera and 200) or (sharded_frame_equals_unsharded and True-True)" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest: rc $? at $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/${tag}_pytest.log)"
timeout 40 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench: rc $? at $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("frames/s %.1f  ms %.3f  e2e %.1f  launches/frame %d  froxel launch ms %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] // d["steps"], d["passes_ms"]["Volumetric light integration"]))
PY
