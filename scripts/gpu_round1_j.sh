#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rel err|Error" gpurun_out/pytest_gpu.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print('tf32', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['inference'], d['gpu_launches'])"
tail -3 gpurun_out/bench_tf32.err
cat > /tmp/infer_only.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
hb = bench.host_batches(1024, 1, 777)
db = bench.to_device(hb[0], dev)
for _ in range(3):
    tr.infer_step(db)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/infer_only.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_infer_tf32.csv python /tmp/infer_only.py > gpurun_out/ncu_infer.log 2>&1
echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edgeconv_fused -s 5 -c 1 -f -o gpurun_out/prof_edgeconv_fused python /tmp/infer_only.py > gpurun_out/ncu_full_fused.log 2>&1
echo "ncu full exit $?"
