#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -s > gpurun_out/pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_tc.log
GNB_PRECISION=tf32 timeout 300 python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/bench_short_tf32.json 2> gpurun_out/bench_short_tf32.err &&
GNB_PRECISION=tf32 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/ncu_tf32.log 2>&1
echo "ncu exit $?"
GNB_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_linear -s 12 -c 2 -o gpurun_out/prof_tc_linear \
    python bench.py --steps 1 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
