#!/bin/bash
# round 2, call aq: re-entry check of HEAD on a fresh box: whole GPU suite, smoke, default bench, reference arm, step breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/aq_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/aq_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aq_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/aq_smoke.log
timeout 900 python bench.py > gpurun_out/aq_bench_1gpu.json 2> gpurun_out/aq_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/aq_bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['e2e'], d['gpu_launches'], d['clocks'], d['roofline']['frac'], d.get('inference',{}).get('value'))"
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/aq_mode_train.log 2>&1; grep -v Warn gpurun_out/aq_mode_train.log | head -40
