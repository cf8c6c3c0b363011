from .layers import DynEdgeConv  # noqa: F401
