"""Helpers of `graphnet.models.utils` that sit on the DynEdge path."""

from typing import Tuple

from torch import Tensor

from graphnet_b200 import ops


def calculate_xyzt_homophily(x: Tensor, edge_index, batch: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Per-event homophily of columns 0..3 (reference: src/graphnet/models/utils.py:13-29).

    Runs the fused global-variables kernel and returns its four homophily
    columns, each `[B, 1]`. `edge_index` may be an int64 `[2, E]` tensor or a
    `KnnGraph`.
    """
    n = x.shape[0]
    graph = edge_index if isinstance(edge_index, ops.KnnGraph) else ops.KnnGraph.from_edge_index(edge_index, n)
    nseg = int(batch.max().item()) + 1 if batch.numel() else 0
    ptr = ops.batch_to_ptr(batch, nseg)
    ones = x.new_ones(nseg)
    g, _ = ops.global_variables(x, graph, ptr, ones)
    f = x.shape[1]
    return g[:, f:f + 1], g[:, f + 1:f + 2], g[:, f + 2:f + 3], g[:, f + 3:f + 4]
