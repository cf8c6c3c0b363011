"""A few inference steps of the bench Trainer (BASELINE configs[1], 1024 events) for ncu: python scripts/r02/infer_only.py PRECISION [STEPS]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import ops  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "f16")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(1024, 1, 777)[0], dev)
for _ in range(steps):
    tr.infer_step(db)
torch.cuda.synchronize()
print("ok")
