#!/bin/bash
# dual-group scatter kernel: tc parity tests first (bounded), then the role probe, then everything
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "dgrad_scatter" > gpurun_out/pytest_scatter.log 2>&1; rc=$?; echo "scatter tests exit $rc"; tail -5 gpurun_out/pytest_scatter.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|timeout\|assert" gpurun_out/pytest_scatter.log | head -20; exit 1; fi
timeout 300 python scripts/linear_probe.py dual > gpurun_out/probe_dual.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe_dual.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference']['value'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['forward_launch']['launch_ms'])"
tail -3 gpurun_out/bench_tf32.err
