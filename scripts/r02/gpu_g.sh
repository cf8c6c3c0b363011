#!/bin/bash
# round 2, call g: kNN two-phase scan, full suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q > gpurun_out/g_pytest_knn.log 2>&1; echo "knn tests exit $?"; tail -4 gpurun_out/g_pytest_knn.log | cut -c1-250
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/g_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -6 gpurun_out/g_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench exit $?"; tail -2 gpurun_out/g_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/g_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "alt", d["alt_precision"]["value"], "launches", d["gpu_launches_per_step"])
    print("timing", d["timing"])
    inf = d["inference"]; print("inference", inf["value"], inf["ms_per_step"], inf.get("e2e"), inf.get("alt_precision"))
    for row in d["kernels"].get("kernels", []):
        if "knn" in row["kernel"] or "pool" in row["kernel"] or "global" in row["kernel"]: print("  ", row)
except Exception as e:
    print("bench unreadable", e)
PY
timeout 300 python bench.py --workload highmult20k --steps 6 --warmup 2 > gpurun_out/g_bench_highmult20k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/g_bench_highmult20k.json')); print('highmult', d['value'], d['ms_per_step'], d['initial_knn_ms'], d['inference'])"
