#!/bin/bash
# first GPU contact: parity tests, smoke, a short bench, and an ncu launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -40 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-inference > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-inference > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
