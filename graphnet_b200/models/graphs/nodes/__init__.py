from .nodes import NodeDefinition, NodesAsPulses  # noqa: F401
