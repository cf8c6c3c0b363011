"""`DynEdgeJINST`: the DynEdge variant of arXiv:2209.03042 (reference: src/graphnet/models/gnn/dynedge_jinst.py:16-152) on the
B200 kernels -- four `DynEdgeConv(aggr="add")` blocks with LeakyReLU MLPs, skip-concatenation, two post-processing Linears,
[max | min | sum | mean] pooling followed by the four homophily values (t, x, y, z) and the raw pulse count, read-out Linear.
Same constructor, attribute names (`conv_add1..4`, `nn1..3`) and therefore `state_dict` keys as the reference."""

from __future__ import annotations

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.components.layers import DynEdgeConv
from graphnet_b200.models.gnn.gnn import GNN


class DynEdgeJINST(GNN):
    def __init__(self, nb_inputs: int, layer_size_scale: int = 4):
        c = layer_size_scale
        l1, l2, l3, l4, l5, l6 = nb_inputs, c * 16 * 2, c * 32 * 2, c * 42 * 2, c * 32 * 2, c * 16 * 2
        super().__init__(nb_inputs, l6)

        def block(n_in: int, hidden: int) -> DynEdgeConv:
            mlp = torch.nn.Sequential(torch.nn.Linear(n_in * 2, hidden), torch.nn.LeakyReLU(),
                                      torch.nn.Linear(hidden, l3), torch.nn.LeakyReLU())
            return DynEdgeConv(mlp, aggr="add", nb_neighbors=8, features_subset=slice(0, 3))
        self.conv_add1 = block(l1, l2)
        self.conv_add2 = block(l3, l4)
        self.conv_add3 = block(l3, l4)
        self.conv_add4 = block(l3, l4)
        self.nn1 = torch.nn.Linear(l3 * 4 + l1, l4)
        self.nn2 = torch.nn.Linear(l4, l5)
        self.nn3 = torch.nn.Linear(4 * l5 + 5, l6)
        self.lrelu = torch.nn.LeakyReLU()

    def forward(self, data) -> Tensor:
        x, batch = data.x, data.batch
        if not x.is_cuda:
            raise RuntimeError("graphnet_b200.DynEdgeJINST runs on CUDA tensors only (no CPU fallback)")
        n_pulses = data.n_pulses
        nseg = int(n_pulses.numel())
        ptr = getattr(data, "ptr", None)
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, nseg)
        graph = data.knn_graph() if hasattr(data, "knn_graph") else None
        if graph is None:
            graph = ops.KnnGraph.from_edge_index(data.edge_index, x.shape[0], 8)
        g, _ = ops.global_variables(x, graph, ptr, n_pulses)          # [mean(x) | h_x h_y h_z h_t | log10 n]: homophily columns
        f = x.shape[1]
        h_x, h_y, h_z, h_t = (g[:, f + i:f + i + 1] for i in range(4))
        x = x.float()
        a, g1 = self.conv_add1.forward_table(x, graph, ptr)
        b, g2 = self.conv_add2.forward_table(a, g1, ptr)
        c, g3 = self.conv_add3.forward_table(b, g2, ptr)
        d, _ = self.conv_add4.forward_table(c, g3, ptr, recompute=False)        # the 4th recomputed graph is never used
        if getattr(self, "_debug_record", False):      # test hook: the graph fed to every block and the block outputs
            self._debug = {"graphs": [graph, g1, g2, g3], "skips": [x, a, b, c, d]}
        x = torch.cat((x, a, b, c, d), dim=1)
        x = self.nn2(self.lrelu(self.nn1(x)))
        pooled = ops.segment_pool(x, ptr, ["max", "min", "sum", "mean"])         # dynedge_jinst.py:125-128 order
        x = torch.cat((pooled, h_t, h_x, h_y, h_z, n_pulses.reshape(-1, 1).to(pooled.dtype)), dim=1)
        x = self.lrelu(x)
        x = self.nn3(x)
        return self.lrelu(x)
