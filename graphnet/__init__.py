"""Import-path shim: the reference's module paths for the symbols of the DynEdge hot path, served by graphnet_b200.

    from graphnet.models.gnn import DynEdge                      # src/graphnet/models/gnn/dynedge.py
    from graphnet.models.components.layers import DynEdgeConv    # src/graphnet/models/components/layers.py
    from graphnet.models.graphs import KNNGraph                  # src/graphnet/models/graphs/graphs.py
    from graphnet.models.graphs.edges import KNNEdges            # src/graphnet/models/graphs/edges/edges.py
    from graphnet.models.detector.icecube import IceCube86       # src/graphnet/models/detector/icecube.py
    from graphnet.models.detector.prometheus import Prometheus   # src/graphnet/models/detector/prometheus.py

so that code and model configs written against the reference (`class_name: DynEdge`, configs/models/*.yml) load unchanged.
Only the hot path's subtree exists: anything else of the reference (datasets, training loops, deployment ...) is out of scope
and raises ImportError as for any missing module. Do not put this directory on the path next to a real GraphNeT install.
"""

import importlib
import sys

_ALIASES = {
    "models": "models", "models.model": "models.model", "models.utils": "models.utils",
    "models.gnn": "models.gnn", "models.gnn.gnn": "models.gnn.gnn", "models.gnn.dynedge": "models.gnn.dynedge",
    "models.gnn.dynedge_jinst": "models.gnn.dynedge_jinst", "models.gnn.particlenet": "models.gnn.particlenet",
    "models.gnn.dynedge_kaggle_tito": "models.gnn.dynedge_kaggle_tito",
    "models.components": "models.components", "models.components.layers": "models.components.layers",
    "models.graphs": "models.graphs", "models.graphs.graphs": "models.graphs.graphs",
    "models.graphs.graph_definition": "models.graphs.graph_definition",
    "models.graphs.edges": "models.graphs.edges", "models.graphs.edges.edges": "models.graphs.edges.edges",
    "models.graphs.nodes": "models.graphs.nodes", "models.graphs.nodes.nodes": "models.graphs.nodes.nodes",
    "models.detector": "models.detector", "models.detector.detector": "models.detector",
    "models.detector.icecube": "models.detector", "models.detector.prometheus": "models.detector",
}
for _ref, _ours in _ALIASES.items():
    sys.modules.setdefault("graphnet." + _ref, importlib.import_module("graphnet_b200." + _ours))
models = sys.modules["graphnet.models"]
