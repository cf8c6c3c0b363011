#!/bin/bash
# tcgen05 bring-up: tensor-core tests first (bounded), then everything, then bench in tf32 mode
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -s > gpurun_out/pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
timeout 300 python -m pytest tests/test_gpu_tc.py -q -s > gpurun_out/pytest_tc_all.log 2>&1; echo "tc exit $?" >> gpurun_out/pytest_tc_all.log
tail -30 gpurun_out/pytest_tc_all.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
GNB_PRECISION=tf32 timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
cat gpurun_out/bench_tf32.json; tail -5 gpurun_out/bench_tf32.err
