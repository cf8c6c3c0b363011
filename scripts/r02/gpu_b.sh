#!/bin/bash
# round 2, call b: 4-stage split kernel, all parity tests on every route, smoke, the reworked bench, launch list
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_tf32x3.py -q -s -x > gpurun_out/b_pytest_x3.log 2>&1; echo "x3 tests exit $?"
grep -E "passed|failed|FAILED|Error|gnb mbar|rel |conversion" gpurun_out/b_pytest_x3.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.log 2>&1; echo "smoke exit $?"; grep -E "smoke\[|Error|assert" gpurun_out/b_smoke.log | head
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench exit $?"; tail -3 gpurun_out/b_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/b_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "alt", d["alt_precision"], "launches", d["gpu_launches_per_step"])
    print("timing", d["timing"]); print("peaks", d["peaks"]); print("whole", d["whole_step"])
    r = d["roofline"]; print("roof", r["kernel"][:40], r["launch_ms"], r["achieved"], r["frac"], "exec", r["executed_frac"])
    for k in ("backward_launch", "single_pass_forward_launch"):
        print(" ", k, r[k]["launch_ms"], r[k]["achieved"], r[k]["frac"])
    inf = d["inference"]; print("inference", inf["value"], inf["ms_per_step"], inf.get("e2e"), inf.get("alt_precision"))
    print("cpu", d["cpu_baseline"])
    kt = d["kernels"]
    if "kernels" in kt:
        print("device us/step", kt["device_us_per_step"])
        for row in kt["kernels"]:
            print("  ", row)
    else:
        print(kt)
except Exception as e:
    print("bench unreadable", e)
PY
timeout 2400 python -m pytest tests -q -m gpu -x -s > gpurun_out/b_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -4 gpurun_out/b_pytest_gpu.log
grep -E "train_step|prometheus50|golden|default config|config #4|tf32 rel|tf32x3 rel" gpurun_out/b_pytest_gpu.log | head -60
timeout 300 python scripts/r02/train_only.py tf32x3 3 > gpurun_out/b_train_only.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/b_launches_x3.csv \
    python scripts/r02/train_only.py tf32x3 3 > gpurun_out/b_ncu.log 2>&1
echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/b_launches_x3.csv 12 > gpurun_out/b_launches_x3_summary.txt 2>&1; head -14 gpurun_out/b_launches_x3_summary.txt
for w in prometheus50 highmult20k percentile16 microbench; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 2 > gpurun_out/b_bench_$w.json 2> gpurun_out/b_bench_$w.err; echo "$w exit $?"; head -c 1500 gpurun_out/b_bench_$w.json; echo; tail -2 gpurun_out/b_bench_$w.err
done
