"""Device-side graph definition: a raw pulse batch that is already on the GPU becomes the DynEdge input there.

Reference path (per event, on the CPU, inside dataloader workers): `GraphDefinition.forward`
(src/graphnet/models/graphs/graph_definition.py:199-247) = dtype cast -> `Detector._standardize`
(detector/detector.py:63-77) -> `NodesAsPulses` (graphs/nodes/nodes.py:123-132) -> `n_pulses` -> `KNNEdges`
(graphs/edges/edges.py:72-80) -> per-feature attributes, then `Batch.from_data_list` (data/dataloader.py:12-18).
Here the same steps run once per BATCH on the device: one standardisation launch over `[N, F]`, `ptr` / `batch`
by prefix sum and one launch, the batched kNN kernel. Results equal the reference's arithmetic (IEEE fp32 subtract /
divide, log10f); no CPU fallback: CPU tensors raise.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import torch

from graphnet_b200 import ops
from graphnet_b200.data import Batch
from graphnet_b200.models.graphs.nodes import NodesAsPulses


def standardize(pulses: torch.Tensor, kinds, subs, divs) -> torch.Tensor:
    """out[:, c] = pulses[:, c], (pulses[:, c] - subs[c]) / divs[c] or log10(pulses[:, c]) for kinds[c] = 0, 1, 2."""
    ops._cuda(pulses)
    if pulses.dtype != torch.float32 or pulses.dim() != 2:
        raise RuntimeError("graphnet_b200.standardize: expected a [N, F] float32 tensor")
    if pulses.stride(1) != 1:
        pulses = pulses.contiguous()
    n, f = pulses.shape
    out = torch.empty(n, f, dtype=torch.float32, device=pulses.device)
    k = (ctypes.c_int32 * f)(*kinds)
    a = (ctypes.c_float * f)(*subs)
    b = (ctypes.c_float * f)(*divs)
    ops._call("gnb_standardize", ops._ptr(pulses), pulses.stride(0), n, f, k, a, b, ops._ptr(out), f, ops._stream())
    return out


def ptr_to_batch(ptr: torch.Tensor, n: int) -> torch.Tensor:
    ops._cuda(ptr)
    batch = torch.empty(n, dtype=torch.int64, device=ptr.device)
    ops._call("gnb_ptr_to_batch", ops._ptr(ptr), ptr.numel() - 1, n, ops._ptr(batch), ops._stream())
    return batch


class DeviceKNNGraph:
    """`KNNGraph` applied to a whole raw batch on the GPU.

    `graph_definition`: a `KNNGraph` (its detector, feature names, k and columns are used; node definition must be
    `NodesAsPulses`). Call with `pulses[N, F]` (raw detector units, events concatenated in order) and `n_pulses[B]`;
    returns the collated `Batch` with `x`, `batch`, `ptr`, `n_pulses`, the per-feature attributes and the kNN table.
    """

    def __init__(self, graph_definition):
        if not isinstance(graph_definition._node_definition, NodesAsPulses):
            raise NotImplementedError("DeviceKNNGraph supports the NodesAsPulses node definition only")
        if graph_definition._perturbation_dict:
            raise NotImplementedError("DeviceKNNGraph does not perturb inputs")
        self._definition = graph_definition
        names = graph_definition._input_feature_names
        self._table = graph_definition._detector.standardisation_table(names)
        self._names = list(graph_definition.output_feature_names)

    def __call__(self, pulses: torch.Tensor, n_pulses: torch.Tensor, ptr: Optional[torch.Tensor] = None) -> Batch:
        ops._cuda(pulses, n_pulses)
        if pulses.shape[1] != len(self._names):
            raise RuntimeError(f"expected {len(self._names)} input features, got {pulses.shape[1]}")
        n = pulses.shape[0]
        if ptr is None:
            ptr = torch.zeros(n_pulses.numel() + 1, dtype=torch.int64, device=pulses.device)
            torch.cumsum(n_pulses.to(torch.int64), 0, out=ptr[1:])
        x = standardize(pulses.to(self._definition.dtype), *self._table)
        graph = Batch(x=x)
        graph.batch = ptr_to_batch(ptr, n)
        graph.ptr = ptr
        graph.n_pulses = n_pulses.to(torch.int32)
        graph = self._definition._edge_definition(graph)
        for idx, name in enumerate(self._names):          # graph_definition.py:243-247
            if name != "x":
                graph[name] = x[:, idx].detach()
        graph["graph_definition"] = self._definition.__class__.__name__
        return graph
