#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-inference > gpurun_out/bench_short_tf32.json 2> gpurun_out/bench_short_tf32.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-inference > gpurun_out/ncu_tf32.log 2>&1
echo "ncu exit $?"
