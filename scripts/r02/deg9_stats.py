"""How many nodes keep k + 1 = 9 neighbours in each of the four graphs of the bench's training batch (the device flag of the
8-slot layout is per layer: one such node switches the layer back to 9 slots), and what the fused forward gains on 8 slots."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphnet_b200 import Data, ops  # noqa: E402
from graphnet_b200.models.graphs.edges import KNNEdges  # noqa: E402

dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1)
ops.set_precision("mixed16")
hb = bench.host_batches(512, 2, 20240607)[0]
db = bench.to_device(hb, dev)
model = tr.backbone
model._debug_record = True
data = tr.edges(tr.make_data(db))
with torch.no_grad():
    model(data)
for li, g in enumerate(model._debug["graphs"]):
    deg = g.deg
    n9 = int((deg == 9).sum())
    print(f"graph {li}: n = {deg.numel()}, nodes with 9 neighbours {n9} ({100.0 * n9 / deg.numel():.3f} %), deg < 8: {int((deg < 8).sum())}")
    f = model._debug["skips"][li][:, :3]
    print("   exact-zero rows in columns 0..2:", int((f == 0).all(1).sum()))
