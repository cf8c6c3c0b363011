from .layers import DynEdgeConv, DynTrans, EdgeConvTito  # noqa: F401
