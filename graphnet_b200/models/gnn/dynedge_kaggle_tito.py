"""`DynEdgeTITO` (reference: src/graphnet/models/gnn/dynedge_kaggle_tito.py:31-278; the TITO solution of the "IceCube --
Neutrinos in Deep Ice" kaggle competition; BASELINE configs[3] names it): `DynTrans` blocks (max-aggregation EdgeConv on
[x_i, x_j - x_i, x_j] with the first Linear hoisted to nodes, residual + LayerNorm, per-event TransformerEncoder) on a STATIC
kNN graph, LeakyReLU post-processing, global pooling, global variables appended after pooling, read-out.
Same constructor, `_conv_layers` / `_post_processing` / `_readout` names and `state_dict` keys as the reference."""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.components.layers import DynTrans
from graphnet_b200.models.gnn.gnn import GNN

GLOBAL_POOLINGS = ("min", "max", "sum", "mean")


class DynEdgeTITO(GNN):
    def __init__(self, nb_inputs: int, features_subset: List[int] = None,
                 dyntrans_layer_sizes: Optional[List[Tuple[int, ...]]] = None, global_pooling_schemes: List[str] = ["max"],
                 use_global_features: bool = True, use_post_processing_layers: bool = True,
                 post_processing_layer_sizes: List[int] = None, readout_layer_sizes: Optional[List[int]] = None,
                 n_head: int = 8, nb_neighbours: int = 8):
        if dyntrans_layer_sizes is None:
            dyntrans_layer_sizes = [(256, 256), (256, 256), (256, 256), (256, 256)]
        assert isinstance(dyntrans_layer_sizes, list) and len(dyntrans_layer_sizes)
        assert all(isinstance(sizes, tuple) and len(sizes) > 0 for sizes in dyntrans_layer_sizes)
        assert all(all(size > 0 for size in sizes) for sizes in dyntrans_layer_sizes)
        if post_processing_layer_sizes is None:
            post_processing_layer_sizes = [336, 256]
        if readout_layer_sizes is None:
            readout_layer_sizes = [256, 128]
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        if isinstance(global_pooling_schemes, list):
            for scheme in global_pooling_schemes:
                assert scheme in GLOBAL_POOLINGS, f"Global pooling scheme {scheme} not supported."
        else:
            assert global_pooling_schemes is None
        assert global_pooling_schemes, ("No global pooling schemes were request, so cannot add global"
                                        " variables after pooling.")
        super().__init__(nb_inputs, readout_layer_sizes[-1])
        self._dyntrans_layer_sizes = dyntrans_layer_sizes
        self._post_processing_layer_sizes = post_processing_layer_sizes
        self._readout_layer_sizes = readout_layer_sizes
        self._global_pooling_schemes = global_pooling_schemes
        self._activation = torch.nn.LeakyReLU()
        self._nb_inputs = nb_inputs
        self._nb_global_variables = 5 + nb_inputs
        self._nb_neighbours = nb_neighbours
        self._features_subset = features_subset or [0, 1, 2, 3]
        self._use_global_features = use_global_features
        self._use_post_processing_layers = use_post_processing_layers
        self._n_head = n_head
        self._construct_layers()

    def _construct_layers(self) -> None:
        self._conv_layers = torch.nn.ModuleList()
        width = self._nb_inputs
        for sizes in self._dyntrans_layer_sizes:
            self._conv_layers.append(DynTrans([width] + list(sizes), aggr="max", features_subset=self._features_subset,
                                              n_head=self._n_head))
            width = sizes[-1]
        if self._use_post_processing_layers:
            layers: List[torch.nn.Module] = []
            dims = [width] + list(self._post_processing_layer_sizes)
            for n_in, n_out in zip(dims[:-1], dims[1:]):
                layers += [torch.nn.Linear(n_in, n_out), self._activation]
            self._post_processing = torch.nn.Sequential(*layers)
            width = dims[-1]
        npool = len(self._global_pooling_schemes) if self._global_pooling_schemes else 1
        width = width * npool + (self._nb_global_variables if self._use_global_features else 0)
        layers = []
        dims = [width] + list(self._readout_layer_sizes)
        for n_in, n_out in zip(dims[:-1], dims[1:]):
            layers += [torch.nn.Linear(n_in, n_out), self._activation]
        self._readout = torch.nn.Sequential(*layers)

    def forward(self, data) -> Tensor:
        x, batch = data.x, data.batch
        if not x.is_cuda:
            raise RuntimeError("graphnet_b200.DynEdgeTITO runs on CUDA tensors only (no CPU fallback)")
        n_pulses = data.n_pulses
        nseg = int(n_pulses.numel())
        ptr = getattr(data, "ptr", None)
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, nseg)
        graph = data.knn_graph() if hasattr(data, "knn_graph") else None
        if graph is None:
            graph = ops.KnnGraph.from_edge_index(data.edge_index, x.shape[0], self._nb_neighbours)
        g = None
        if self._use_global_features:       # [mean(x) | h_x h_y h_z h_t | log10 n_pulses]: dynedge_kaggle_tito.py:229-250
            g, _ = ops.global_variables(x, graph, ptr, n_pulses)
        x = x.float()
        for conv in self._conv_layers:      # static graph (dynedge_kaggle_tito.py:264-265)
            x = conv.forward_table(x, graph, ptr)
        if self._use_post_processing_layers:
            x = self._post_processing(x)
        x = ops.segment_pool(x, ptr, self._global_pooling_schemes)
        if self._use_global_features:
            x = torch.cat([x, g], dim=1)
        return self._readout(x)
