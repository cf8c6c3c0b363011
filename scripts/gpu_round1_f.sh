#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|tf32 rel|FAILED|Error" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_fp32.json')); print('fp32', d['value'], d['ms_per_step'], d['inference'], d['gpu_launches'])"
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench tf32 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['inference'], d['gpu_launches'], d['roofline'])"
tail -3 gpurun_out/bench_tf32.err
timeout 300 python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/bench_short_tf32.json 2> gpurun_out/bench_short_tf32.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/ncu_tf32.log 2>&1
echo "ncu exit $?"
