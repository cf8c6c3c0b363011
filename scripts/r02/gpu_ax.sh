#!/bin/bash
# round 2, call ax: f16 inference on the fused forward (resident W2 plane); whole GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/ax_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ax_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ax_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/ax_smoke.log
timeout 900 python bench.py > gpurun_out/ax_bench_1gpu.json 2> gpurun_out/ax_bench.err; echo "bench exit $?"; tail -3 gpurun_out/ax_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ax_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], d["inference"].get("e2e"), "launches", d.get("gpu_launches_per_step"))
r = d["roofline"]; print(r["bound"], r["kernel"][:50], r["launch_ms"], r["achieved"], r["peak"], r["frac"], r.get("tensor_view"))
ri = d["inference"]["roofline"]; print("inf", ri["bound"], ri["kernel"][:50], ri["launch_ms"], ri["achieved"], ri["frac"], ri.get("tensor_view"))
for k in d["kernels"]["kernels"][:16]: print(k)
PY
