"""Where a DynEdgeTITO training step spends its time (CUPTI kernel table + host-side wall): python scripts/r02/tito_profile.py"""
import collections
import re
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import bench_workloads as bw  # noqa: E402
from graphnet_b200 import ops  # noqa: E402
from graphnet_b200.models.gnn import DynEdgeTITO  # noqa: E402
from graphnet_b200.synthetic import event_sizes, make_batch  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
dev = torch.device("cuda", 0)
nev = 256
sizes = event_sizes(nev, np.random.default_rng(5), sigma=1.0, n_max=256)
raw = make_batch(nev, seed=5, sizes=sizes)
db = {k: torch.from_numpy(np.ascontiguousarray(raw[k])).to(dev) for k in ("x", "batch", "n_pulses", "energy")}
torch.manual_seed(0)
tr = bw.EnergyTrainer(dev, 7, backbone=DynEdgeTITO(7, global_pooling_schemes=bench.POOLS))
for what, step in (("train", tr.train_step), ("infer", tr.infer_step)):
    for _ in range(4):
        step(db)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step(db)
    t_host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 5
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step(db)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", ev.name)
            name = re.sub(r"\(.*", "", name)[:70]
            agg[name][0] += 1
            agg[name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"== {what}: host enqueue {t_host * 1e3:.2f} ms, wall {t_all * 1e3:.2f} ms / step, device busy {tot / 2e3:.2f} ms, "
          f"{sum(v[0] for v in agg.values()) / 2:.0f} launches")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"   {t / 2:9.1f} us {100 * t / tot:5.1f}% n={c / 2:5.1f} {k}")
