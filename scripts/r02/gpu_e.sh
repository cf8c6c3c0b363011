#!/bin/bash
# round 2, call e: integer-only splitter, SIMT tail fixes; full suite + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tf32x3.py tests/test_gpu_ops.py -q -s > gpurun_out/e_pytest_x3.log 2>&1; echo "x3+ops tests exit $?"
grep -E "passed|failed|FAILED|gnb mbar|tf32x3 linear|aggregating GEMM|max grad|inference rel" gpurun_out/e_pytest_x3.log | cut -c1-200 | head -24
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench exit $?"; tail -2 gpurun_out/e_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/e_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "alt", d["alt_precision"]["value"], "launches", d["gpu_launches_per_step"])
    print("timing", d["timing"]); print("whole", d["whole_step"])
    r = d["roofline"]; print("roof", r["launch_ms"], r["achieved"], r["frac"], "exec", r["executed_frac"])
    inf = d["inference"]; print("inference", inf["value"], inf["ms_per_step"], inf.get("e2e"), inf.get("alt_precision"))
    for row in d["kernels"].get("kernels", []):
        print("  ", row)
except Exception as e:
    print("bench unreadable", e)
PY
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/e_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -8 gpurun_out/e_pytest_gpu.log | cut -c1-300
timeout 300 python scripts/r02/train_only.py tf32x3 3 > gpurun_out/e_train_only.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/e_launches_x3.csv \
    python scripts/r02/train_only.py tf32x3 3 > gpurun_out/e_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/e_launches_x3.csv 20 > gpurun_out/e_launches_x3_summary.txt 2>&1; head -22 gpurun_out/e_launches_x3_summary.txt
