"""Runs only the dominant kernel (tcgen05 Linear on the [E_rows, 336] x [336, 256] edge GEMM) for ncu --set full."""
import sys, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(512, 1, 20240607)[0], dev)
data = tr.edges(tr.make_data(db))
graph = data.knn_graph()
rows = graph.n * graph.width
lin = tr.backbone._conv_layers[1].nn[2]
h = ops._round_pad(torch.rand(rows, lin.in_features, device=dev))
packed = ops._tc_pack_weight(lin.weight.detach(), (0,), (lin.in_features,))
for _ in range(4):
    ops._tc_linear((h,), packed, lin.bias.detach(), lin.out_features, ops.ACT_RELU, round_out=False)
torch.cuda.synchronize()
print("rows", rows)
