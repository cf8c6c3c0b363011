"""`ParticleNeT` (reference: src/graphnet/models/gnn/particlenet.py:21-255, after arXiv:1902.08570) on the B200 kernels:
`DynEdgeConv(aggr="mean")` blocks whose MLPs are Linear [BatchNorm1d] activation chains of any depth (generic EdgeConv
route: gather kernel -> the user's `nn` -> mean-aggregation kernel), optional kNN recompute after every block (`dynamic`),
global pooling, read-out with dropout. Same constructor, `_conv_layers` / `_readout` names and `state_dict` keys."""

from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.components.layers import DynEdgeConv
from graphnet_b200.models.gnn.gnn import GNN

GLOBAL_POOLINGS = ("min", "max", "sum", "mean")


class ParticleNeT(GNN):
    def __init__(self, nb_inputs: int, *, nb_neighbours: int = 16,
                 features_subset: Optional[Union[List[int], slice]] = None, dynamic: bool = True,
                 dynedge_layer_sizes: Optional[List[Tuple[int, ...]]] = None,
                 readout_layer_sizes: Optional[List[int]] = None,
                 global_pooling_schemes: Optional[Union[str, List[str]]] = "mean",
                 activation_layer: Optional[str] = "relu", add_batchnorm_layer: bool = True,
                 dropout_readout: float = 0.1, skip_readout: bool = False):
        if features_subset is None:
            features_subset = slice(0, 3)
        if dynedge_layer_sizes is None:
            dynedge_layer_sizes = [(64, 64, 64), (128, 128, 128), (256, 256, 256)]
        sizes_checked = [tuple(s) if isinstance(s, list) else s for s in dynedge_layer_sizes]
        assert len(sizes_checked) and all(isinstance(s, tuple) and len(s) > 0 for s in sizes_checked)
        assert all(all(size > 0 for size in s) for s in sizes_checked)
        if readout_layer_sizes is None:
            readout_layer_sizes = [256]
        assert isinstance(readout_layer_sizes, list) and len(readout_layer_sizes) and all(s > 0 for s in readout_layer_sizes)
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        if isinstance(global_pooling_schemes, list):
            for scheme in global_pooling_schemes:
                assert scheme in GLOBAL_POOLINGS, f"Global pooling scheme {scheme} not supported."
        else:
            assert global_pooling_schemes is None
        if activation_layer is None or activation_layer.lower() == "relu":
            activation = torch.nn.ReLU()
        elif activation_layer.lower() == "gelu":
            activation = torch.nn.GELU()
        else:
            raise ValueError(f"Activation layer {activation_layer} not supported.")
        super().__init__(nb_inputs, readout_layer_sizes[-1])
        self._dynedge_layer_sizes = sizes_checked
        self._readout_layer_sizes = readout_layer_sizes
        self._global_pooling_schemes = global_pooling_schemes
        self._activation = activation
        self._nb_inputs = nb_inputs
        self._nb_neighbours = nb_neighbours
        self._features_subset = features_subset
        self._dynamic = dynamic
        self._add_batchnorm_layer = add_batchnorm_layer
        self._dropout_readout = dropout_readout
        self._skip_readout = skip_readout
        self._construct_layers()

    def _construct_layers(self) -> None:
        self._conv_layers = torch.nn.ModuleList()
        width = self._nb_inputs
        for sizes in self._dynedge_layer_sizes:
            layers: List[torch.nn.Module] = []
            dims = [width] + list(sizes)
            for ix, (n_in, n_out) in enumerate(zip(dims[:-1], dims[1:])):
                layers.append(torch.nn.Linear(2 * n_in if ix == 0 else n_in, n_out))
                if self._add_batchnorm_layer:
                    layers.append(torch.nn.BatchNorm1d(n_out))
                layers.append(self._activation)
            self._conv_layers.append(DynEdgeConv(torch.nn.Sequential(*layers), aggr="mean", nb_neighbors=self._nb_neighbours,
                                                 features_subset=self._features_subset))
            width = dims[-1]
        npool = len(self._global_pooling_schemes) if self._global_pooling_schemes else 1
        dims = [width * npool] + list(self._readout_layer_sizes)
        readout: List[torch.nn.Module] = []
        for n_in, n_out in zip(dims[:-1], dims[1:]):
            readout += [torch.nn.Linear(n_in, n_out), self._activation, torch.nn.Dropout(self._dropout_readout)]
        self._readout = torch.nn.Sequential(*readout)

    def _global_pooling(self, x: Tensor, batch: Tensor = None, ptr: Tensor = None) -> Tensor:
        assert self._global_pooling_schemes
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, int(batch.max().item()) + 1)
        return ops.segment_pool(x, ptr, self._global_pooling_schemes)

    def forward(self, data) -> Tensor:
        x, batch = data.x, data.batch
        if not x.is_cuda:
            raise RuntimeError("graphnet_b200.ParticleNeT runs on CUDA tensors only (no CPU fallback)")
        ptr = getattr(data, "ptr", None)
        if ptr is None:
            n_pulses = getattr(data, "n_pulses", None)
            nseg = int(n_pulses.numel()) if n_pulses is not None else int(batch.max().item()) + 1
            ptr = ops.batch_to_ptr(batch, nseg)
        graph = data.knn_graph() if hasattr(data, "knn_graph") else None
        if graph is None:
            graph = ops.KnnGraph.from_edge_index(data.edge_index, x.shape[0], self._nb_neighbours)
        x = x.float()
        last = len(self._conv_layers) - 1
        graphs, outs = [graph], [x]
        for li, conv in enumerate(self._conv_layers):
            # the graph recomputed after the last block (and every recompute when not `dynamic`) is never used
            x, new_graph = conv.forward_table(x, graph, ptr, recompute=self._dynamic and li < last)
            if new_graph is not None:
                graph = new_graph
            graphs.append(graph)
            outs.append(x)
        if getattr(self, "_debug_record", False):      # test hook: the graph fed to every block and the block outputs
            self._debug = {"graphs": graphs[:-1], "skips": outs}
        if not self._skip_readout:
            if self._global_pooling_schemes:
                x = self._global_pooling(x, ptr=ptr)
            x = self._readout(x)
        return x
