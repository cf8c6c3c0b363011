#!/bin/bash
# round 2, call a: the split-operand (tf32x3) kernels -- parity tests, bench in both tensor-core modes, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_tf32x3.py -q -s -x > gpurun_out/a_pytest_x3.log 2>&1; echo "x3 tests exit $?"
grep -E "passed|failed|FAILED|Error|gnb mbar|rel " gpurun_out/a_pytest_x3.log | head -40
timeout 600 python bench.py --precision tf32x3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_x3.json 2> gpurun_out/a_bench_x3.err; echo "bench x3 exit $?"
timeout 600 python bench.py --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_tf32.json 2> gpurun_out/a_bench_tf32.err; echo "bench tf32 exit $?"
python - <<'PY'
import json
for f in ("a_bench_x3", "a_bench_tf32"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"], "launches", d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/a_bench_x3.err
timeout 300 python scripts/r02/train_only.py tf32x3 3 > gpurun_out/a_train_only.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/a_launches_x3.csv \
    python scripts/r02/train_only.py tf32x3 3 > gpurun_out/a_ncu.log 2>&1
echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/a_launches_x3.csv 24 > gpurun_out/a_launches_x3_summary.txt 2>&1; head -30 gpurun_out/a_launches_x3_summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -5 gpurun_out/a_pytest_gpu.log
