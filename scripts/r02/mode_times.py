"""Step time of the bench Trainer per precision mode + per-kernel device time (CUPTI via torch.profiler):
python scripts/r02/mode_times.py MODE[,MODE...] [train|infer]"""
import collections
import re
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import ops  # noqa: E402

modes = (sys.argv[1] if len(sys.argv) > 1 else "tf32x3,bf16x3,bf16,tf32").split(",")
what = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
nev = 512 if what == "train" else 1024
dbs = [bench.to_device(hb, dev) for hb in bench.host_batches(nev, 2, 20240607 if what == "train" else 777)]
step = tr.train_step if what == "train" else tr.infer_step
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for mode in modes:
    tag = mode
    ops.STORE_DZ = "+dz" in mode
    ops.UNFUSED_FORWARD = "+unf" in mode
    mode = mode.replace("+dz", "").replace("+unf", "")
    ops.set_precision(mode)
    for i in range(6):
        step(dbs[i & 1])
    torch.cuda.synchronize()
    ts = []
    for i in range(12):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(dbs[i & 1])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"== {tag} {what}: median {med:.3f} ms / step = {nev / med * 1e3:.0f} events/s (min {ts[0]:.3f})", flush=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for i in range(2):
            step(dbs[i & 1])
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", ev.name)
            name = re.sub(r"\(.*", "", name)[:60]
            agg[name][0] += 1
            agg[name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"   device time {tot / 2:.0f} us / step")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {t / 2:9.1f} us {100 * t / tot:5.1f}% n={c / 2:5.1f} {k}")
