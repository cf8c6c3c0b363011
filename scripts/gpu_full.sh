#!/bin/bash
# full GPU check: parity tests, bench (tf32 default), optional launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['frac'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_tf32.err
