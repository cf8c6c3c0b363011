"""The dominant kind::f16 launches of the mixed16 training step, alone, for `ncu --set full`:
python scripts/r02/profile_f16_kernels.py   (3 rounds of: scattering data gradient, aggregating forward, weight gradient)"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import ops  # noqa: E402

ops.set_precision("mixed16")
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
d = bench.dominant_launches(tr, db)
for _ in range(3):
    d["dgrad_scatter_f16"]()
    d["agg_fwd_f16x3"]()
    d["wgrad_f16"]()
torch.cuda.synchronize()
print("rows", d["rows"], "n", d["n"], "edges", d["edges"])
