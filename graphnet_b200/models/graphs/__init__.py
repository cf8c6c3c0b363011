from .graph_definition import GraphDefinition  # noqa: F401
from .graphs import KNNGraph  # noqa: F401
