#!/bin/bash
# round 2, 2-GPU check: the driver's launch line for N = 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/az_bench_2gpu.json 2> gpurun_out/az_bench_2gpu.err; echo "bench2 exit $?"; tail -3 gpurun_out/az_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/az_bench_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], d["n_gpus"], d.get("timing", {}).get("per_rank_step_ms_median"))
PY
