from .gnn import GNN  # noqa: F401
from .dynedge import DynEdge  # noqa: F401
