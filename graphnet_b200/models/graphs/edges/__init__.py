from .edges import EdgeDefinition, KNNEdges  # noqa: F401
