"""Node definitions (reference: src/graphnet/models/graphs/nodes/nodes.py:17-132)."""

from __future__ import annotations

from typing import List, Tuple

import torch

from graphnet_b200.data import Data
from graphnet_b200.models.model import Model


class NodeDefinition(Model):
    def __init__(self, input_feature_names: List[str] = None) -> None:
        super().__init__()
        if input_feature_names is not None:
            self.set_output_feature_names(input_feature_names)

    def set_output_feature_names(self, input_feature_names: List[str]) -> None:
        self._output_feature_names = self._define_output_feature_names(input_feature_names)

    def set_number_of_inputs(self, input_feature_names: List[str]) -> None:
        self.nb_inputs = len(input_feature_names)

    @property
    def nb_outputs(self) -> int:
        return len(self._output_feature_names)

    def forward(self, x: torch.Tensor) -> Tuple[Data, List[str]]:
        return self._construct_nodes(x), self._output_feature_names

    def _define_output_feature_names(self, input_feature_names: List[str]) -> List[str]:
        raise NotImplementedError

    def _construct_nodes(self, x: torch.Tensor) -> Data:
        raise NotImplementedError


class NodesAsPulses(NodeDefinition):
    """Every pulse is a node."""

    def _define_output_feature_names(self, input_feature_names: List[str]) -> List[str]:
        return input_feature_names

    def _construct_nodes(self, x: torch.Tensor) -> Data:
        return Data(x=x)
