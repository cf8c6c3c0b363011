"""`GNN` base class (reference: src/graphnet/models/gnn/gnn.py:11-35)."""

from abc import abstractmethod

from torch import Tensor

from graphnet_b200.models.model import Model


class GNN(Model):
    """Interface shared by graph-network backbones: input / output widths and `forward(data)`."""

    def __init__(self, nb_inputs: int, nb_outputs: int) -> None:
        super().__init__()
        self._nb_inputs = nb_inputs
        self._nb_outputs = nb_outputs

    @property
    def nb_inputs(self) -> int:
        return self._nb_inputs

    @property
    def nb_outputs(self) -> int:
        return self._nb_outputs

    @abstractmethod
    def forward(self, data) -> Tensor:
        """Map a batch of event graphs to `[B or N, nb_outputs]`."""
