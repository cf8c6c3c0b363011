#!/bin/bash
set -x
mkdir -p gpurun_out
GNB_WGRAD_SWAP=0 timeout 300 python -m pytest tests/test_gpu_tc.py -q -k wgrad > gpurun_out/wgrad_swap0.log 2>&1; echo "swap0 exit $?" >> gpurun_out/wgrad_swap0.log
tail -4 gpurun_out/wgrad_swap0.log
GNB_WGRAD_SWAP=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q -k wgrad > gpurun_out/wgrad_swap1.log 2>&1; echo "swap1 exit $?" >> gpurun_out/wgrad_swap1.log
tail -4 gpurun_out/wgrad_swap1.log
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|tf32 rel" gpurun_out/pytest_gpu.log
GNB_PRECISION=tf32 timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
cat gpurun_out/bench_tf32.json; tail -5 gpurun_out/bench_tf32.err
