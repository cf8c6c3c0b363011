#!/bin/bash
# round 2, 8-GPU check: the driver's launch line for N = 8 (and N = 4)
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bb_bench_${n}gpu.json 2> gpurun_out/bb_bench_${n}gpu.err; echo "bench$n exit $?"; tail -2 gpurun_out/bb_bench_${n}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bb_bench_${n}gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "inf", d["inference"]["value"], d["n_gpus"], d.get("timing", {}).get("per_rank_step_ms_median"))
PY
done
