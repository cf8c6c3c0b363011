"""The dz-free backward launches alone (for ncu / timing): python scripts/r02/profile_masked.py"""
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
for name in ("wgrad_f16", "wgrad_f16_masked", "dgrad_scatter_f16", "dgrad_scatter_f16_masked", "agg_fwd_f16x3"):
    print(name, round(bench._time_launch(d[name]) * 1e6, 1), "us", flush=True)

for flags in (1, 2, 3):
    ops._call("gnb_wgrad_set_debug", flags)
    print("wgrad_f16_masked dbg", flags, round(bench._time_launch(d["wgrad_f16_masked"]) * 1e6, 1), "us", flush=True)
ops._call("gnb_wgrad_set_debug", 0)
