from .gnn import GNN  # noqa: F401
from .dynedge import DynEdge  # noqa: F401
from .dynedge_jinst import DynEdgeJINST  # noqa: F401
from .particlenet import ParticleNeT  # noqa: F401
from .dynedge_kaggle_tito import DynEdgeTITO  # noqa: F401
