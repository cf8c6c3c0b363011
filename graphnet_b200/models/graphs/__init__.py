from .graph_definition import GraphDefinition  # noqa: F401
from .graphs import KNNGraph  # noqa: F401
from .device import DeviceKNNGraph  # noqa: F401
