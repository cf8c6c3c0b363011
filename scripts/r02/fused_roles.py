"""Fused EdgeConv forward launch with parts of the builders switched off (timing only): python scripts/r02/fused_roles.py"""
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
for name in ("fused_fwd_f16x3", "fused_fwd_f16x3_inference"):
    for flags, what in ((0, "full"), (8, "no P gathers"), (16, "no Q gathers"), (24, "no gathers")):
        ops._call("gnb_linear_set_debug", flags)
        print(name, what, round(bench._time_launch(d[name]) * 1e6, 1), "us", flush=True)
ops._call("gnb_linear_set_debug", 0)
for name in ("wgrad_f16_masked", "dgrad_scatter_f16_masked", "agg_fwd_f16x3"):
    print(name, round(bench._time_launch(d[name]) * 1e6, 1), "us", flush=True)
