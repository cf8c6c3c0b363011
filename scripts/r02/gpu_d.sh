#!/bin/bash
# round 2, call d: bf16 correction products in the split kernel; every parity test with kernel-side ReLU decisions
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python -m pytest tests/test_gpu_tf32x3.py -q -s > gpurun_out/d_pytest_x3_$i.log 2>&1; echo "x3 tests run $i exit $?"
grep -E "passed|failed|FAILED|gnb mbar|tf32x3 linear|aggregating GEMM|max grad|inference rel" gpurun_out/d_pytest_x3_$i.log | cut -c1-260 | head -24
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; echo "smoke exit $?"; grep -E "smoke\[|Error|assert" gpurun_out/d_smoke.log | head
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/d_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -12 gpurun_out/d_pytest_gpu.log | cut -c1-300
grep -E "train_step|prometheus50|golden |default config|config #4|^\.?[a-z_]+ (fp32|tf32x3): out" gpurun_out/d_pytest_gpu.log | cut -c1-260 | head -70
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench exit $?"; tail -2 gpurun_out/d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/d_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "alt", d["alt_precision"]["value"], "launches", d["gpu_launches_per_step"])
    print("timing", d["timing"]); print("whole", d["whole_step"])
    r = d["roofline"]; print("roof", r["launch_ms"], r["achieved"], r["frac"], "exec", r["executed_frac"])
    inf = d["inference"]; print("inference", inf["value"], inf["ms_per_step"], inf.get("e2e"), inf.get("alt_precision"))
    for row in d["kernels"].get("kernels", []):
        print("  ", row)
except Exception as e:
    print("bench unreadable", e)
PY
timeout 300 python scripts/r02/train_only.py tf32x3 3 > gpurun_out/d_train_only.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/d_launches_x3.csv \
    python scripts/r02/train_only.py tf32x3 3 > gpurun_out/d_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/d_launches_x3.csv 8 > gpurun_out/d_launches_x3_summary.txt 2>&1; head -10 gpurun_out/d_launches_x3_summary.txt
