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


class PercentileClusters(NodeDefinition):
    """One node per cluster of pulses that agree on `cluster_on` (e.g. one node per DOM); every other feature is summarised
    by the given percentiles of its values inside the cluster, optionally followed by log10(cluster size).

    Reference: `PercentileClusters` (src/graphnet/models/graphs/nodes/nodes.py:135-217) over
    `cluster_summarize_with_percentiles` / `gather_cluster_sequence` (src/graphnet/models/graphs/utils.py:32-172).
    Host-side, per event, like the reference (it runs inside dataloader workers); the arithmetic is restated without
    the NaN-padded `[clusters, max multiplicity]` detour: sort once, find the cluster boundaries, and take
    `np.percentile` per group of equally sized clusters (float64, numpy's default linear interpolation -- the same
    routine `np.nanpercentile` ends up calling on each row's non-NaN part). Cluster order = `np.lexsort` over the
    cluster columns (last column most significant), as in the reference.
    """

    def __init__(self, cluster_on: List[str], percentiles: List[int], add_counts: bool = True,
                 input_feature_names: List[str] = None) -> None:
        self._cluster_on = list(cluster_on)
        self._percentiles = list(percentiles)
        self._add_counts = add_counts
        super().__init__(input_feature_names=input_feature_names)

    def _define_output_feature_names(self, input_feature_names: List[str]) -> List[str]:
        self._cluster_indices = [input_feature_names.index(name) for name in self._cluster_on]
        summarised = [name for name in input_feature_names if name not in self._cluster_on]
        self._summarization_indices = [input_feature_names.index(name) for name in summarised]
        names = list(self._cluster_on)
        names += [f"{name}_pct{pct}" for name in summarised for pct in self._percentiles]
        if self._add_counts:
            names.append("counts")
        return names

    def _construct_nodes(self, x: torch.Tensor) -> Data:
        import numpy as np
        if not hasattr(self, "_summarization_indices"):
            raise AttributeError(f"{self.__class__.__name__} needs `input_feature_names` (constructor or GraphDefinition)")
        a = x.numpy()
        order = np.lexsort(tuple(a[:, c] for c in self._cluster_indices))
        a = a[order]
        keys = a[:, self._cluster_indices]
        first = np.ones(a.shape[0], dtype=bool)
        first[1:] = np.any(keys[1:] != keys[:-1], axis=1)
        starts = np.flatnonzero(first)
        counts = np.diff(np.append(starts, a.shape[0]))
        n_clusters, n_pct = starts.size, len(self._percentiles)
        out = np.empty((n_clusters, len(self._cluster_indices) + n_pct * len(self._summarization_indices)
                        + (1 if self._add_counts else 0)), dtype=np.float64)
        out[:, :len(self._cluster_indices)] = keys[starts]
        col = len(self._cluster_indices)
        sizes = np.unique(counts)
        for f in self._summarization_indices:
            vals = a[:, f].astype(np.float64)
            for m in sizes:                                   # all clusters with m pulses at once
                sel = np.flatnonzero(counts == m)
                mat = vals[starts[sel, None] + np.arange(m)[None, :]]
                out[sel, col:col + n_pct] = np.percentile(mat, self._percentiles, axis=1).T
            col += n_pct
        if self._add_counts:
            out[:, col] = np.log10(counts)
        return Data(x=torch.tensor(out))
