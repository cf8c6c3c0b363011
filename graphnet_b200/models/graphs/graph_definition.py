"""`GraphDefinition`: raw pulse array -> `Data` (reference: graphs/graph_definition.py:148-248).

Only the steps that feed the DynEdge path are kept: dtype cast, detector standardisation, node
definition, `n_pulses`, edge definition, per-feature attributes and the definition stamp. Sensor
masking, inactive-sensor padding, perturbation and truth/label attachment are dataloader concerns
outside the hot path; simple truth dictionaries are still attached so
collated batches carry their labels.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from graphnet_b200.data import Data
from graphnet_b200.models.model import Model


class GraphDefinition(Model):
    def __init__(self, detector, node_definition=None, edge_definition=None,
                 input_feature_names: Optional[List[str]] = None, dtype: Optional[torch.dtype] = torch.float,
                 perturbation_dict: Optional[Dict[str, float]] = None, seed: Any = None,
                 add_inactive_sensors: bool = False, sensor_mask: Optional[List[int]] = None,
                 string_mask: Optional[List[int]] = None, sort_by: Optional[str] = None, repeat_labels: bool = False,
                 **kwargs: Any):
        super().__init__()
        # Reference options that belong to the dataloader side of the graph definition (graph_definition.py:33-37,
        # 184-207) and are NOT implemented here: refuse them instead of silently building a different graph.
        dropped = {"add_inactive_sensors": add_inactive_sensors, "sensor_mask": sensor_mask, "string_mask": string_mask,
                   "sort_by": sort_by, "repeat_labels": repeat_labels}
        used = [k for k, v in dropped.items() if v not in (None, False)]
        if used:
            raise NotImplementedError(
                f"graphnet_b200.GraphDefinition does not implement {used} (outside the DynEdge hot path); "
                "apply them in the dataloader before the graph definition")
        from graphnet_b200.models.graphs.nodes import NodesAsPulses
        self._detector = detector
        self._node_definition = node_definition or NodesAsPulses()
        self._edge_definition = edge_definition
        self.dtype = dtype
        self._perturbation_dict = perturbation_dict
        self.rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        if input_feature_names is None:
            input_feature_names = list(getattr(detector, "feature_map", lambda: {})().keys())
        self._input_feature_names = list(input_feature_names)
        self._node_definition.set_number_of_inputs(self._input_feature_names)
        self._node_definition.set_output_feature_names(self._input_feature_names)
        self.nb_inputs = len(self._input_feature_names)
        self.nb_outputs = self._node_definition.nb_outputs
        self.output_feature_names = self._node_definition._output_feature_names

    def forward(self, input_features: np.ndarray, input_feature_names: List[str],
                truth_dicts: Optional[List[Dict[str, Any]]] = None, custom_label_functions=None,
                loss_weight_column: Optional[str] = None, loss_weight: Optional[float] = None,
                loss_weight_default_value: Optional[float] = None, data_path: Optional[str] = None) -> Data:
        assert input_features.shape[1] == len(input_feature_names)
        if self._perturbation_dict:
            input_features = np.array(input_features, dtype=np.float64, copy=True)
            for name, std in self._perturbation_dict.items():
                idx = input_feature_names.index(name)
                input_features[:, idx] = self.rng.normal(input_features[:, idx], std)
        feats = torch.tensor(input_features, dtype=self.dtype)
        feats = self._detector(feats, input_feature_names)
        graph, node_feature_names = self._node_definition(feats)
        graph.x = graph.x.type(self.dtype)
        graph.n_pulses = torch.tensor(len(input_features), dtype=torch.int32)
        if self._edge_definition is not None:
            graph = self._edge_definition(graph)
        if data_path is not None:
            graph["dataset_path"] = data_path
        if truth_dicts is not None:
            for truth in truth_dicts:
                for key, val in truth.items():
                    try:
                        graph[key] = torch.tensor(val)
                    except Exception:
                        pass
        if custom_label_functions is not None:
            for key, fn in custom_label_functions.items():
                graph[key] = fn(graph)
        for idx, name in enumerate(node_feature_names):
            if name != "x":
                graph[name] = graph.x[:, idx].detach()
        graph["graph_definition"] = self.__class__.__name__
        return graph
