"""`DynEdgeConv`: EdgeConv followed by a kNN recompute in the layer's output space.

Reference: src/graphnet/models/components/layers.py:20-69 (a PyG `EdgeConv`
whose `forward` also calls `knn_graph` on `x[:, features_subset]`).

B200 design. The per-edge MLP input `[x_i, x_j - x_i]` is never materialised
when the MLP has the DynEdge shape `Linear, ReLU, Linear, ReLU`: the first
Linear is hoisted to nodes (`W1 [x_i; x_j-x_i] + b1 = (W1a-W1b) x_i + b1 + W1b x_j`),
so the only per-edge dense contraction left is the second Linear, and the
k-neighbour aggregation reads each message once. Any other `nn` (LayerNorm,
GELU, BatchNorm, 3-layer MLPs ...) runs through the generic route: gather
kernel -> the user's `nn` -> aggregation kernel. Both routes are CUDA only.
"""

from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.model import Model


def _hoistable(nn: Any) -> bool:
    if not isinstance(nn, torch.nn.Sequential) or len(nn) != 4:
        return False
    return (isinstance(nn[0], torch.nn.Linear) and isinstance(nn[1], torch.nn.ReLU)
            and isinstance(nn[2], torch.nn.Linear) and isinstance(nn[3], torch.nn.ReLU))


def _has_batch_statistics(nn: Any) -> bool:
    return isinstance(nn, torch.nn.Module) and any(
        isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in nn.modules())


def pad_columns(t: Tensor, mult: int) -> Tensor:
    pad = (-t.shape[1]) % mult
    return t if pad == 0 else torch.nn.functional.pad(t, (0, pad))


class DynEdgeConv(Model):
    """Dynamical edge convolution layer."""

    def __init__(
        self,
        nn: Callable,
        aggr: str = "max",
        nb_neighbors: int = 8,
        features_subset: Optional[Union[Sequence[int], slice]] = None,
        **kwargs: Any,
    ):
        if features_subset is None:
            features_subset = slice(None)
        assert isinstance(features_subset, (list, slice))
        assert aggr in ops.AGGR, f"aggr={aggr!r} not supported"
        super().__init__()
        self.nn = nn
        self.aggr = aggr
        self.nb_neighbors = nb_neighbors
        self.features_subset = features_subset

    # -- EdgeConv: out_i = AGG_j nn([x_i, x_j - x_i]) --------------------------------------------
    def edge_conv(self, x: Tensor, graph: ops.KnnGraph) -> Tensor:
        if _hoistable(self.nn):
            lin1, lin2 = self.nn[0], self.nn[2]
            c = lin1.in_features // 2
            cp = x.shape[1]                       # x may carry zero pad columns beyond c
            w1 = lin1.weight
            wa, wb = w1[:, :c], w1[:, c:]
            wcat = torch.cat([wa - wb, wb], dim=0)                       # [2H, c]
            if cp != c:
                wcat = torch.nn.functional.pad(wcat, (0, cp - c))
            wcat = pad_columns(wcat, 4)
            if wcat.shape[1] != cp:
                x = pad_columns(x, 4)
            bcat = None
            if lin1.bias is not None:
                bcat = torch.cat([lin1.bias, torch.zeros_like(lin1.bias)])
            pq = ops.linear_act(x, wcat, bcat, ops.ACT_NONE, round_out=False)   # [N, 2H] = [P | Q], added in fp32
            if lin1.out_features % 4 == 0 and lin2.out_features % 4 == 0 and self.aggr in ("add", "sum", "mean"):
                return ops.edgeconv_hoisted(pq, lin2.weight, lin2.bias, graph, self.aggr)
            if lin1.out_features % 4 == 0:
                h = ops.edge_hidden(pq, graph, ops.ACT_RELU)             # [N*W, H]
                m = ops.linear_act(h, lin2.weight, lin2.bias, ops.ACT_RELU)
                return ops.edge_aggregate(m, graph, self.aggr)
        # generic route: any callable `nn`
        u = ops.edge_cat(x, graph)
        if _has_batch_statistics(self.nn):
            valid = (torch.arange(graph.width, device=x.device).unsqueeze(0) < graph.deg.unsqueeze(1)).flatten()
            mv = self.nn(u[valid])
            m = mv.new_zeros(u.shape[0], mv.shape[1])
            m[valid] = mv
        else:
            m = self.nn(u)
        return ops.edge_aggregate(m.float(), graph, self.aggr)

    def recompute_graph(self, x: Tensor, ptr: Tensor) -> ops.KnnGraph:
        cols = ops.resolve_columns(self.features_subset, x.shape[1])
        return ops.knn_table(x, cols, ptr, self.nb_neighbors)

    def forward_table(self, x: Tensor, graph: ops.KnnGraph, ptr: Tensor,
                      recompute: bool = True) -> Tuple[Tensor, Optional[ops.KnnGraph]]:
        x = self.edge_conv(x, graph)
        return x, (self.recompute_graph(x, ptr) if recompute else None)

    def forward(self, x: Tensor, edge_index, batch: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Reference signature: returns the new features and the recomputed `edge_index`."""
        n = x.shape[0]
        graph = edge_index if isinstance(edge_index, ops.KnnGraph) else ops.KnnGraph.from_edge_index(edge_index, n)
        if batch is None:
            ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
        else:
            ptr = ops.batch_to_ptr(batch, int(batch.max().item()) + 1)
        x, new_graph = self.forward_table(x, graph, ptr)
        return x, new_graph.edge_index()
