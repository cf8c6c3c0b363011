#!/bin/bash
# round 2, call f (2 GPUs): overlapped all-reduce + cost-model sharding; reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/f_bench_2gpu.json 2> gpurun_out/f_bench_2gpu.err; echo "bench 2gpu exit $?"; tail -3 gpurun_out/f_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap --no-inference --no-alt-precision > gpurun_out/f_bench_2gpu_no_overlap.json 2> gpurun_out/f_bench_2gpu_no_overlap.err; echo "bench 2gpu no-overlap exit $?"
python - <<'PY'
import json
for f in ("f_bench_2gpu", "f_bench_2gpu_no_overlap"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "timing", d["timing"])
        print("   inference", (d.get("inference") or {}).get("value"), "alt", (d.get("alt_precision") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/f_bench_ref_2gpu.json 2> gpurun_out/f_bench_ref_2gpu.err; echo "ref arm exit $?"; cat gpurun_out/f_bench_ref_2gpu.json | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_dynedge.py tests/test_gpu_users.py -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/f_pytest.log | cut -c1-250
