"""Runs only the two dominant launches of the training step (gemm_tc_pair_kernel: scattering data-gradient GEMM and
aggregating forward GEMM of an EdgeConv layer) for `ncu --set full`:  python scripts/profile_top_kernel.py"""
import sys, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
d = bench.dominant_launches(tr, db)
for _ in range(3):
    d["dgrad_scatter"]()
    d["agg_fwd"]()
torch.cuda.synchronize()
print("rows", d["rows"], "n", d["n"], "edges", d["edges"])
