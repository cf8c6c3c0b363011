import sys, torch
sys.path.insert(0, ".")
from graphnet_b200 import Data, ops
from graphnet_b200.models.gnn import DynEdge
from graphnet_b200.models.graphs.edges import KNNEdges
from graphnet_b200.synthetic import make_batch
ops.set_precision("mixed16")
raw = make_batch(24, seed=11, n_max=500)
x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
torch.manual_seed(3)
model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
model._debug_record = True
data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
res = {}
for name, unf, grad in (("fused_train", False, True), ("unfused_train", True, True), ("fused_inf", False, False), ("unfused_inf", True, False)):
    ops.UNFUSED_FORWARD = unf
    with torch.set_grad_enabled(grad):
        y = model(data)
    res[name] = (y.detach().clone(), [s.detach().clone() for s in model._debug["skips"]])
ops.UNFUSED_FORWARD = False
ref = res["unfused_train"]
for k, (y, sk) in res.items():
    print(k, "out", float((y - ref[0]).abs().max() / ref[0].abs().max()), "skips", [float((a - b).abs().max() / b.abs().max().clamp(min=1e-9)) for a, b in zip(sk, ref[1])])
