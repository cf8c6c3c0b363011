"""kNN kernel alone on the bench batches (512 / 1024 events): python scripts/knn_probe.py [events]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from graphnet_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
nev = int(sys.argv[1]) if len(sys.argv) > 1 else 512
db = bench.to_device(bench.host_batches(nev, 1, 20240607)[0], dev)
x = db["x"]
ptr = ops.batch_to_ptr(db["batch"], nev)
n = x.shape[0]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return b.elapsed_time(e) / reps * 1e3


sizes = (ptr[1:] - ptr[:-1]).cpu()
print(f"events {nev} nodes {n} max event {int(sizes.max())} sum n^2 {int((sizes.double() ** 2).sum())}", flush=True)
for variant in (0, 1):
    lib.gnb_knn_set_variant(variant)
    t = timed(lambda: ops.knn_table(x, [0, 1, 2], ptr, 8))
    print(f"kNN variant {variant}: {t:.0f} us", flush=True)
lib.gnb_knn_set_variant(0)
xl = torch.randn(n, 256, device=dev)
t = timed(lambda: ops.knn_table(xl, [0, 1, 2], ptr, 8))
print(f"kNN on a 256-wide latent tensor (random normal): {t:.0f} us", flush=True)
