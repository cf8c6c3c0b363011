"""Edge definitions (reference: src/graphnet/models/graphs/edges/edges.py:14-80)."""

from __future__ import annotations

from typing import List

import torch

from graphnet_b200 import ops
from graphnet_b200.models.model import Model


class EdgeDefinition(Model):
    def forward(self, graph):
        return self._construct_edges(graph)

    def _construct_edges(self, graph):
        raise NotImplementedError


class KNNEdges(EdgeDefinition):
    """Connect every node to its `nb_nearest_neighbours` nearest nodes of the same event.

    On a CUDA `Data`/`Batch` this runs the batched kNN kernel and attaches the neighbour table
    (`graph.edge_index` then materialises lazily). On a CPU graph -- the situation inside the
    reference's dataloader workers (dataset.py:642-651), where CUDA cannot be used -- the edges are
    *deferred*: call this definition again on the collated device batch (the reference's own
    `KNNEdges` already accepts `graph.batch`, edges.py:74-78). There is no CPU kNN in this package.
    """

    def __init__(self, nb_nearest_neighbours: int, columns: List[int] = [0, 1, 2]):
        super().__init__()
        self._nb_nearest_neighbours = nb_nearest_neighbours
        self._columns = list(columns)

    def _construct_edges(self, graph):
        x = graph.x
        if not x.is_cuda:
            graph.edge_index = None      # deferred to the device batch
            return graph
        n = x.shape[0]
        batch = getattr(graph, "batch", None)
        ptr = getattr(graph, "ptr", None)
        if ptr is None:
            if batch is None:
                ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
            else:
                n_pulses = getattr(graph, "n_pulses", None)
                nseg = int(n_pulses.numel()) if n_pulses is not None and n_pulses.dim() else int(batch.max().item()) + 1
                ptr = ops.batch_to_ptr(batch, nseg)
                try:                      # a collated Batch carries `ptr` (torch_geometric sets it): keep it for the model, which
                    graph.ptr = ptr       # would otherwise rebuild it from `batch` with a second launch
                except (AttributeError, TypeError):
                    pass
        table = ops.knn_table(x, self._columns, ptr, self._nb_nearest_neighbours)
        if hasattr(graph, "set_knn_graph"):
            graph.set_knn_graph(table)
        else:
            graph.edge_index = table.edge_index()
        return graph
