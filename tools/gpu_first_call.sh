#!/bin/bash
# First GPU call of a round: everything that was written without a GPU, in one gpurun call.
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Outputs land in gpurun_out/ (merged back by gpurun). Each step has its own timeout so that a hang costs one step, not the box.
set -u
mkdir -p gpurun_out
export PLAIN_TEST_UNVERIFIED=1
timeout 600 python -m pytest tests/test_zz_single_pass_gpu.py -q -x -s > gpurun_out/first_staged_single_pass.log 2>&1; echo "staged single-pass: $?"
timeout 600 python -m pytest tests/test_zz_fast_contract_gpu.py -q -s > gpurun_out/first_staged_fast_contract.log 2>&1; echo "staged fast contract: $?"
unset PLAIN_TEST_UNVERIFIED
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/first_pytest_gpu.log 2>&1; echo "pytest -m gpu: $?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/first_bench_exact.json 2> gpurun_out/first_bench_exact.err; echo "bench exact: $?"
timeout 300 python bench.py --no-cpu-baseline --contract fast > gpurun_out/first_bench_fast.json 2> gpurun_out/first_bench_fast.err; echo "bench fast: $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/first_launches_fast.csv python bench.py --no-cpu-baseline --contract fast --steps 2 --warmup 3 --no-graph > gpurun_out/first_ncu_fast.log 2>&1; echo "ncu launch list (fast): $?"
tail -3 gpurun_out/first_staged_single_pass.log gpurun_out/first_staged_fast_contract.log gpurun_out/first_pytest_gpu.log
cat gpurun_out/first_bench_exact.json gpurun_out/first_bench_fast.json | cut -c1-400
