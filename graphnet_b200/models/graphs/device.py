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
from graphnet_b200.models.graphs.nodes import NodesAsPulses, PercentileClusters


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


def percentile_clusters(x: torch.Tensor, ptr: torch.Tensor, cluster_indices, summarization_indices, percentiles,
                        add_counts: bool = True):
    """`PercentileClusters` (reference: graphs/nodes/nodes.py:135-217 over graphs/utils.py:32-172, per event on the CPU) for a
    whole standardised pulse batch on the device: nodes [M, len(cluster) + P * len(summarised) (+ 1)] fp32 and the node `ptr`.

    Same arithmetic as the reference's numpy route: clusters = runs of equal `cluster_indices` columns inside an event after a
    lexicographic sort (last cluster column most significant, np.lexsort), each summarised feature by numpy's default
    percentile (virtual index (m - 1) q / 100, linear interpolation a + (b - a) t, evaluated as b - (b - a)(1 - t) for
    t >= 0.5) in float64 on the float32 values, log10(count) in float64, everything cast to the graph dtype at the end.
    Sorting is segmented by composition: stable sorts by the least significant key first, the event index last."""
    ops._cuda(x, ptr)
    n = x.shape[0]
    dev = x.device
    nseg = ptr.numel() - 1
    batch = ptr_to_batch(ptr, n)
    order = torch.arange(n, device=dev)
    for c in cluster_indices:                               # np.lexsort: the LAST key is the primary one
        order = order[torch.sort(x[order, c], stable=True).indices]
    order = order[torch.sort(batch[order], stable=True).indices]
    xs, bs = x[order], batch[order]
    keys = xs[:, list(cluster_indices)]
    first = torch.ones(n, dtype=torch.bool, device=dev)
    if n > 1:
        first[1:] = (keys[1:] != keys[:-1]).any(dim=1) | (bs[1:] != bs[:-1])
    cid = torch.cumsum(first.to(torch.int64), 0) - 1         # cluster of every (sorted) pulse
    starts = torch.nonzero(first).flatten()
    m = int(starts.numel())                                   # (one host sync: the node count sizes the output)
    counts = torch.diff(torch.cat([starts, torch.tensor([n], device=dev)]))
    n_pct = len(percentiles)
    width = len(cluster_indices) + n_pct * len(summarization_indices) + (1 if add_counts else 0)
    out = torch.empty(m, width, dtype=torch.float64, device=dev)
    out[:, :len(cluster_indices)] = keys[starts].double()
    cm1 = (counts - 1).double()
    col = len(cluster_indices)
    for f in summarization_indices:
        v = xs[:, f]
        o2 = torch.sort(v, stable=True).indices
        o2 = o2[torch.sort(cid[o2], stable=True).indices]     # ascending values inside every cluster
        vals = v[o2].double()
        for q in percentiles:
            vi = cm1 * (q / 100.0)                             # numpy: (n - 1) * quantile
            lo = torch.floor(vi)
            t = vi - lo
            lo_i = lo.to(torch.int64)
            hi_i = torch.minimum(lo_i + 1, counts - 1)
            a, b = vals[starts + lo_i], vals[starts + hi_i]
            diff = b - a
            r = a + diff * t
            r = torch.where(t >= 0.5, b - diff * (1.0 - t), r)
            out[:, col] = r
            col += 1
    if add_counts:
        out[:, col] = torch.log10(counts.double())
    node_ptr = torch.zeros(nseg + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(bs[starts], minlength=nseg), 0, out=node_ptr[1:])
    return out, node_ptr


class DeviceKNNGraph:
    """`KNNGraph` applied to a whole raw batch on the GPU.

    `graph_definition`: a `KNNGraph` (its detector, feature names, k and columns are used; node definition `NodesAsPulses`
    or `PercentileClusters`). Call with `pulses[N, F]` (raw detector units, events concatenated in order) and `n_pulses[B]`;
    returns the collated `Batch` with `x`, `batch`, `ptr`, `n_pulses` (the RAW pulse counts, graph_definition.py:213), the
    per-feature attributes and the kNN table.
    """

    def __init__(self, graph_definition):
        node_def = graph_definition._node_definition
        if not isinstance(node_def, (NodesAsPulses, PercentileClusters)):
            raise NotImplementedError("DeviceKNNGraph supports the NodesAsPulses and PercentileClusters node definitions only")
        if graph_definition._perturbation_dict:
            raise NotImplementedError("DeviceKNNGraph does not perturb inputs")
        self._definition = graph_definition
        self._clusters = node_def if isinstance(node_def, PercentileClusters) else None
        names = graph_definition._input_feature_names
        self._n_inputs = len(names)
        self._table = graph_definition._detector.standardisation_table(names)
        self._names = list(graph_definition.output_feature_names)

    def __call__(self, pulses: torch.Tensor, n_pulses: torch.Tensor, ptr: Optional[torch.Tensor] = None) -> Batch:
        ops._cuda(pulses, n_pulses)
        if pulses.shape[1] != self._n_inputs:
            raise RuntimeError(f"expected {self._n_inputs} input features, got {pulses.shape[1]}")
        n = pulses.shape[0]
        if ptr is None:
            ptr = torch.zeros(n_pulses.numel() + 1, dtype=torch.int64, device=pulses.device)
            torch.cumsum(n_pulses.to(torch.int64), 0, out=ptr[1:])
        x = standardize(pulses.to(self._definition.dtype), *self._table)
        if self._clusters is not None:
            c = self._clusters
            nodes, ptr = percentile_clusters(x, ptr, c._cluster_indices, c._summarization_indices, c._percentiles, c._add_counts)
            x = nodes.to(self._definition.dtype)
            n = x.shape[0]
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
