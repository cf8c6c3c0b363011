"""The max-aggregating GEMM launch alone at the bench's layer size (79 k pulses, 256 -> 256, k = 8), forward only, for ncu and
timing against the unfused route: python scripts/r02/aggmax_only.py [PRECISION] [REPS]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import ops  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
graph = tr.edges(tr.make_data(db)).knn_graph()
hid = c_out = 256
g = torch.Generator(device="cuda").manual_seed(0)
pq = torch.randn(graph.n, 2 * hid, device=dev, generator=g)
w2 = torch.randn(c_out, hid, device=dev, generator=g) / hid ** 0.5
b2 = torch.randn(c_out, device=dev, generator=g) * 0.1


def fused():
    return ops.edgeconv_hoisted_max(pq, w2, b2, graph, ops.ACT_LEAKY, ops.ACT_LEAKY)


def unfused():
    a1 = ops.edge_hidden(pq, graph, ops.ACT_NONE)
    m = torch.nn.functional.leaky_relu(torch.nn.functional.linear(torch.nn.functional.leaky_relu(a1), w2, b2))
    return ops.edge_aggregate(m, graph, "max")


with torch.no_grad():
    for _ in range(reps):
        y = fused()
    torch.cuda.synchronize()
    t_f = bench._time_launch(fused)
    t_u = bench._time_launch(unfused)
    err = float((fused() - unfused()).abs().max() / unfused().abs().max())
print(f"n {graph.n} rows {graph.n * graph.width}: max-epilogue route {t_f * 1e6:.1f} us, unfused route {t_u * 1e6:.1f} us, "
      f"difference {err:.2e}")
