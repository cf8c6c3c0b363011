#!/bin/bash
# round 2, call q: HEAD after the container re-creation -- full GPU test suite, default bench (all legs), launch list of the mixed16 step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/q_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -5 gpurun_out/q_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/q_bench_1gpu.json 2> gpurun_out/q_bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/q_bench_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/q_bench_1gpu.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d.get("inference", {}).get("value"), "launches", d.get("gpu_launches_per_step"))
print(json.dumps(d.get("roofline"))[:600])
PY
timeout 300 python scripts/r02/train_only.py mixed16 3 > gpurun_out/q_train_only.log 2>&1; echo "train_only exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_train_mixed16.csv \
   python scripts/r02/train_only.py mixed16 3 > gpurun_out/q_ncu_list.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/q_launches_train_mixed16.csv 30 > gpurun_out/q_launches_train_mixed16_summary.txt 2>&1; head -34 gpurun_out/q_launches_train_mixed16_summary.txt
timeout 300 python scripts/r02/mode_times.py mixed16 infer > gpurun_out/q_mode_infer.log 2>&1; cat gpurun_out/q_mode_infer.log | grep -v Warn
