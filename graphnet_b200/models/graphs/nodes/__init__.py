from .nodes import NodeDefinition, NodesAsPulses, PercentileClusters  # noqa: F401
