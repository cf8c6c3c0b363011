"""graphnet_b200: B200-native (sm_100a) implementation of GraphNeT's DynEdge hot path.

Module layout mirrors the reference's import paths for the path's public symbols:
`graphnet_b200.models.gnn.DynEdge`, `graphnet_b200.models.components.layers.DynEdgeConv`,
`graphnet_b200.models.graphs.KNNGraph`, `graphnet_b200.models.graphs.edges.KNNEdges`.
"""

from .data import Batch, Data  # noqa: F401

__all__ = ["Batch", "Data"]
