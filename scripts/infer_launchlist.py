"""Inference only (BASELINE configs[1], 1024 events): a few steps for an ncu launch list."""
import sys, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
hb = bench.host_batches(1024, 2, seed0=777)
db = [bench.to_device(h, dev) for h in hb]
for i in range(4):
    tr.infer_step(db[i % 2])
torch.cuda.synchronize()
print("done")
