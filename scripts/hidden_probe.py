"""Time the per-node hidden-layer forward kernel and the kNN kernel alone on training-step shapes."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from graphnet_b200 import ops, _lib
ops.set_precision("tf32")
lib = _lib.load()
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
data = tr.edges(tr.make_data(db))
graph = data.knn_graph()
n = graph.n


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


for hid in (336, 128):
    pq = torch.randn(n, 2 * hid, device=dev)
    h = torch.empty(n * 9, hid, device=dev)
    mld = 4 * ((hid + 127) // 128)
    hmask = torch.empty((n + 13) // 14 * 126, mld, dtype=torch.int32, device=dev)
    t = timed(lambda: ops._call("gnb_edge_hidden_fwd_mask", ops._ptr(pq), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg),
                                9, n, 1 | 0x100, ops._ptr(h), hid, ops._ptr(hmask), mld, ops._stream()))
    print(f"hidden fwd (mask) hid={hid}: {t:.0f} us, {4.0 * n * 9 * hid / t / 1e6:.2f} TB/s written", flush=True)
x = db["x"]
ptr = data.ptr if hasattr(data, "ptr") else None
ptr = ops.batch_to_ptr(db["batch"], 512)
for variant in (0, 1):
    lib.gnb_knn_set_variant(variant)
    t = timed(lambda: ops.knn_table(x, [0, 1, 2], ptr, 8))
    print(f"kNN variant {variant}: {t:.0f} us", flush=True)
lib.gnb_knn_set_variant(0)
xl = torch.randn(n, 256, device=dev)
t = timed(lambda: ops.knn_table(xl, [0, 1, 2], ptr, 8))
print(f"kNN on a 256-wide latent tensor (random normal): {t:.0f} us", flush=True)
