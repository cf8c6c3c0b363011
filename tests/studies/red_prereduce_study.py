"""CPU study for DESIGN.md section 8 item 3: how many of the scattering epilogue's fp32 L2 reductions would disappear if the dQ
contributions of a tile were first summed per distinct source node in shared memory? For the graphs of a DynEdge forward
(initial kNN on xyz + the three latent-space graphs, random-init weights) on synthetic IceCube-like events, count edges vs
distinct (tile, source) pairs for tiles of 14 and 28 consecutive target nodes (the sub-tile / tile of the kernels).
Lives under tests/ because it drives the oracle: python tests/studies/red_prereduce_study.py [events]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from graphnet_b200.synthetic import make_batch  # noqa: E402
from helpers import namespace  # noqa: E402
import oracle.dynedge_oracle as orc  # noqa: E402


def main():
    nev = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    raw = make_batch(nev, seed=20240607)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    ptr = orc.batch_to_ptr(batch)
    ei0 = orc.knn_graph_ref(x[:, :3], 8, ptr=ptr)
    torch.manual_seed(0)
    model = orc.DynEdgeRef(7, global_pooling_schemes=["min", "max", "mean", "sum"])
    with torch.no_grad():
        _, inter = model(namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses), return_intermediates=True)
    print(f"events {nev}, pulses {x.shape[0]}, median pulses/event {int(np.median(raw['n_pulses']))}")
    for li, ei in enumerate(inter["graphs"][:4]):
        src, dst = ei[0].numpy().astype(np.int64), ei[1].numpy().astype(np.int64)
        line = f"graph of conv layer {li + 1}: {src.size} edges"
        for tile in (14, 28):
            pairs = np.unique((dst // tile) * (1 << 32) + src)
            line += f" | tile of {tile} nodes: {pairs.size} distinct (tile, source) pairs = {pairs.size / src.size:.2f} of the reductions"
        print(line)


if __name__ == "__main__":
    main()
