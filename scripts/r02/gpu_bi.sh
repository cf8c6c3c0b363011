#!/bin/bash
# round 2, last rehearsal of the driver's round-end sequence on the final tree: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/bi_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/bi_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bi_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/bi_smoke.log
timeout 900 python bench.py > gpurun_out/bi_bench_1gpu.json 2> gpurun_out/bi_bench_1gpu.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bi_bench_reference.json 2> gpurun_out/bi_bench_reference.err; echo "reference exit $?"; cut -c1-200 gpurun_out/bi_bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bi_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], d["inference"].get("e2e", {}).get("value"), "launches", d.get("gpu_launches_per_step"), d["clocks"])
for k in d["kernels"]["kernels"][:12]: print(k["kernel"], k["us_per_step"], k.get("frac"), k.get("hbm_frac"))
PY
