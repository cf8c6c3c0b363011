#!/bin/bash
# round 2, call ad: gradient absmax folded into the producing GEMM epilogue; ptr kept on the batch: parity + timings; default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_train_step.py tests/test_gpu_dynedge.py tests/test_gpu_config0.py -q -x > gpurun_out/ad_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ad_pytest.log
timeout 300 python scripts/r02/mode_times.py mixed16 train > gpurun_out/ad_mode_train.log 2>&1; grep -v Warn gpurun_out/ad_mode_train.log | head -24
timeout 300 python scripts/r02/gaps.py mixed16 2>&1 | grep -v Warn | head -16
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/ad_bench_1gpu.json 2> gpurun_out/ad_bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/ad_bench_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/ad_bench_1gpu.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d.get("inference", {}).get("value"), d.get("inference", {}).get("precision"), "launches", d.get("gpu_launches_per_step"))
r = d["roofline"]; print(r["kernel"][:60], r["launch_ms"], r["achieved"], r["frac"], r["executed_frac"])
for k in ("backward_launch", "weight_gradient_launch"):
    print(k, r[k]["launch_ms"], r[k]["achieved"], r[k]["frac"])
print([ (a["precision"][:8], a["value"]) for a in d["alt_precision"]])
PY
