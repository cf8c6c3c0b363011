#!/bin/bash
# parity tests + bench (1 GPU): bash scripts/gpu_r1_m.sh [pytest -k expr]
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['forward_launch']['launch_ms'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_tf32.err
