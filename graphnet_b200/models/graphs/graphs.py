"""`KNNGraph` (reference: src/graphnet/models/graphs/graphs.py:13-58)."""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from graphnet_b200.models.graphs.edges import KNNEdges
from graphnet_b200.models.graphs.graph_definition import GraphDefinition


class KNNGraph(GraphDefinition):
    """Graph representation whose edges go to the k nearest neighbours in `columns`."""

    def __init__(self, detector, node_definition=None, input_feature_names: Optional[List[str]] = None,
                 dtype: Optional[torch.dtype] = torch.float, perturbation_dict: Optional[Dict[str, float]] = None,
                 seed: Any = None, nb_nearest_neighbours: int = 8, columns: List[int] = [0, 1, 2],
                 **kwargs: Any) -> None:
        super().__init__(detector=detector, node_definition=node_definition,
                         edge_definition=KNNEdges(nb_nearest_neighbours=nb_nearest_neighbours, columns=columns),
                         dtype=dtype, input_feature_names=input_feature_names,
                         perturbation_dict=perturbation_dict, seed=seed, **kwargs)

    def build_edges(self, batch):
        """Run the (deferred) edge definition on a collated device batch."""
        return self._edge_definition(batch)
