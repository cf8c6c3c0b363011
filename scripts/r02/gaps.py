"""Idle time between the kernels of one training step (CUPTI timestamps via torch.profiler): python scripts/r02/gaps.py [MODE]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import ops  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "mixed16")
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
for _ in range(5):
    tr.train_step(db)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_step(db)
    torch.cuda.synchronize()
ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name], key=lambda t: t[0])
span = ev[-1][1] - ev[0][0]
busy = sum(e[1] - e[0] for e in ev)
gaps = [(ev[i + 1][0] - ev[i][1], ev[i][2][:50], ev[i + 1][2][:50]) for i in range(len(ev) - 1)]
print(f"kernels {len(ev)} span {span:.0f} us busy {busy:.0f} us idle {span - busy:.0f} us ({100 * (span - busy) / span:.1f} %)")
gs = sorted(g[0] for g in gaps)
print("gap us: median", gs[len(gs) // 2], "p90", gs[int(0.9 * len(gs))], "max", gs[-1], "sum", sum(gs))
for g in sorted(gaps, key=lambda g: -g[0])[:12]:
    print(f"  {g[0]:7.1f} us  {g[1]} -> {g[2]}")
# where the idle time sits: gaps by the kernel that FOLLOWS them (launch ramp of that kernel), last step only
import collections
third = [g for g in gaps[2 * len(gaps) // 3:]]
by = collections.defaultdict(lambda: [0, 0.0])
for g in third:
    by[g[2]][0] += 1
    by[g[2]][1] += g[0]
print("last step: gaps", len(third), "sum", round(sum(g[0] for g in third), 1), "us")
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {t:7.1f} us over {c:3d} gaps (mean {t / c:5.2f})  before {k}")
