#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|tf32 rel|FAILED|Error|error" gpurun_out/pytest_gpu.log | head -20
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
for prec in tf32 fp32; do
timeout 600 python bench.py --steps 5 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; echo "bench $prec exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$prec.json')); print('$prec', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference'], d['gpu_launches'])"
tail -3 gpurun_out/bench_$prec.err
done
timeout 300 python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/bench_short_tf32.json 2> gpurun_out/bench_short_tf32.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_tf32.csv \
    python bench.py --steps 2 --warmup 1 --precision tf32 --no-cpu-baseline --no-inference > gpurun_out/ncu_tf32.log 2>&1
echo "ncu exit $?"
