#!/bin/bash
# Round 2, first GPU call: the cases staged at the end of round 1, the fast contract's one measurement, and the
# "before" ncu captures (launch list + --set full of every kernel above 1 % of the frame) of the round-1 build.
set -u
mkdir -p gpurun_out
export PLAIN_TEST_UNVERIFIED=1
timeout 900 python -m pytest tests/test_zz_single_pass_gpu.py -q -s > gpurun_out/r2a_staged_single_pass.log 2>&1; echo "staged single-pass: $?"
timeout 600 python -m pytest tests/test_zz_fast_contract_gpu.py -q -s > gpurun_out/r2a_staged_fast_contract.log 2>&1; echo "staged fast contract: $?"
unset PLAIN_TEST_UNVERIFIED
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_exact.json 2> gpurun_out/r2a_bench_exact.err; echo "bench exact: $?"
timeout 300 python bench.py --no-cpu-baseline --contract fast > gpurun_out/r2a_bench_fast.json 2> gpurun_out/r2a_bench_fast.err; echo "bench fast: $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 --no-graph > gpurun_out/r2a_ncu_launches.log 2>&1; echo "ncu launch list: $?"
K='regex:sdfDiffuseTrace|giSpatialFilter|temporalFilterKernel|gbufferShading|bloomUpsample|bloomDownsample|froxel|volum|giUpscale|giTemporal|tonemapping|applyBloom'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 69 -c 23 -f -o gpurun_out/r2a_before python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > gpurun_out/r2a_ncu_full.log 2>&1; echo "ncu full: $?"
ls -la gpurun_out/r2a_before.ncu-rep
tail -3 gpurun_out/r2a_staged_single_pass.log gpurun_out/r2a_staged_fast_contract.log
cat gpurun_out/r2a_bench_exact.json gpurun_out/r2a_bench_fast.json | cut -c1-300
