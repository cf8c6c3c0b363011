"""DynEdge on B200: same module API, constructor and `state_dict` keys as the reference
(src/graphnet/models/gnn/dynedge.py:21-349), CUDA kernels underneath.

Differences in *how* (never in *what*):
  * global variables are gathered per node (`g[batch]`) by the kernel that computes them instead of
    the dense `[N,B]` mask / `[N,B,G]` product of dynedge.py:308-317 (bit-identical result);
  * the skip-concatenation of dynedge.py:328 is never materialised: the first post-processing
    Linear runs as a K-split GEMM over the per-layer outputs;
  * graphs travel between layers as fixed-width neighbour tables (no int64 `[2,E]` tensors, no
    host synchronisation), and the kNN recompute after the last conv -- whose result the
    reference discards (dynedge.py:323-325) -- is skipped.
"""

from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.components.layers import DynEdgeConv, _hoistable, pad_columns
from graphnet_b200.models.gnn.gnn import GNN

GLOBAL_POOLINGS = ("min", "max", "sum", "mean")
_DEFAULT_CONV_SIZES = [(128, 256), (336, 256), (336, 256), (336, 256)]


def _mlp(widths: List[int], act: torch.nn.Module, norm: bool, double_first: bool = False) -> torch.nn.Sequential:
    mods: List[torch.nn.Module] = []
    for ix in range(len(widths) - 1):
        fan_in = widths[ix] * (2 if (double_first and ix == 0) else 1)
        mods.append(torch.nn.Linear(fan_in, widths[ix + 1]))
        if norm:
            mods.append(torch.nn.LayerNorm(widths[ix + 1]))
        mods.append(act)
    return torch.nn.Sequential(*mods)


def _is_linear_relu_chain(seq: torch.nn.Sequential) -> bool:
    mods = list(seq)
    if len(mods) == 0 or len(mods) % 2:
        return False
    return all(isinstance(mods[i], torch.nn.Linear) and isinstance(mods[i + 1], torch.nn.ReLU)
               for i in range(0, len(mods), 2))


class DynEdge(GNN):
    """DynEdge (dynamical edge convolutional) model."""

    def __init__(
        self,
        nb_inputs: int,
        *,
        nb_neighbours: int = 8,
        features_subset: Optional[Union[List[int], slice]] = None,
        dynedge_layer_sizes: Optional[List[Tuple[int, ...]]] = None,
        post_processing_layer_sizes: Optional[List[int]] = None,
        readout_layer_sizes: Optional[List[int]] = None,
        global_pooling_schemes: Optional[Union[str, List[str]]] = None,
        add_global_variables_after_pooling: bool = False,
        activation_layer: Optional[str] = None,
        add_norm_layer: bool = False,
        skip_readout: bool = False,
    ):
        if features_subset is None:
            features_subset = slice(0, 3)
        if dynedge_layer_sizes is None:
            dynedge_layer_sizes = list(_DEFAULT_CONV_SIZES)
        assert isinstance(dynedge_layer_sizes, list) and len(dynedge_layer_sizes)
        assert all(isinstance(s, tuple) and len(s) > 0 and all(w > 0 for w in s) for s in dynedge_layer_sizes)
        if post_processing_layer_sizes is None:
            post_processing_layer_sizes = [336, 256]
        assert isinstance(post_processing_layer_sizes, list) and len(post_processing_layer_sizes)
        assert all(w > 0 for w in post_processing_layer_sizes)
        if readout_layer_sizes is None:
            readout_layer_sizes = [128]
        assert isinstance(readout_layer_sizes, list) and len(readout_layer_sizes)
        assert all(w > 0 for w in readout_layer_sizes)
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        if isinstance(global_pooling_schemes, list):
            for scheme in global_pooling_schemes:
                assert scheme in GLOBAL_POOLINGS, f"Global pooling scheme {scheme} not supported."
        else:
            assert global_pooling_schemes is None
        if add_global_variables_after_pooling:
            assert global_pooling_schemes, (
                "No global pooling schemes were request, so cannot add global variables after pooling.")
        if activation_layer is None or activation_layer.lower() == "relu":
            activation = torch.nn.ReLU()
        elif activation_layer.lower() == "gelu":
            activation = torch.nn.GELU()
        else:
            raise ValueError(f"Activation layer {activation_layer} not supported.")

        super().__init__(nb_inputs, readout_layer_sizes[-1])

        self._dynedge_layer_sizes = dynedge_layer_sizes
        self._post_processing_layer_sizes = post_processing_layer_sizes
        self._readout_layer_sizes = readout_layer_sizes
        self._global_pooling_schemes = global_pooling_schemes
        self._add_global_variables_after_pooling = add_global_variables_after_pooling
        self._activation = activation
        self._nb_inputs = nb_inputs
        self._nb_global_variables = 5 + nb_inputs
        self._nb_neighbours = nb_neighbours
        self._features_subset = features_subset
        self._add_norm_layer = add_norm_layer
        self._skip_readout = skip_readout
        self._construct_layers()

    def _construct_layers(self) -> None:
        node_width = self._nb_inputs
        if not self._add_global_variables_after_pooling:
            node_width += self._nb_global_variables
        self._conv_layers = torch.nn.ModuleList()
        width = node_width
        for sizes in self._dynedge_layer_sizes:
            mlp = _mlp([width] + list(sizes), self._activation, self._add_norm_layer, double_first=True)
            self._conv_layers.append(DynEdgeConv(mlp, aggr="add", nb_neighbors=self._nb_neighbours,
                                                 features_subset=self._features_subset))
            width = sizes[-1]
        skip_width = node_width + sum(s[-1] for s in self._dynedge_layer_sizes)
        self._post_processing = _mlp([skip_width] + list(self._post_processing_layer_sizes),
                                     self._activation, self._add_norm_layer)
        npool = len(self._global_pooling_schemes) if self._global_pooling_schemes else 1
        readout_in = self._post_processing_layer_sizes[-1] * npool
        if self._add_global_variables_after_pooling:
            readout_in += self._nb_global_variables
        self._readout = _mlp([readout_in] + list(self._readout_layer_sizes), self._activation, False)

    # -- pieces kept under the reference's method names -----------------------------------------
    def _global_pooling(self, x: Tensor, batch: Tensor = None, ptr: Tensor = None) -> Tensor:
        assert self._global_pooling_schemes
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, int(batch.max().item()) + 1)
        return ops.segment_pool(x, ptr, self._global_pooling_schemes)

    def _calculate_global_variables(self, x: Tensor, graph: ops.KnnGraph, ptr: Tensor, n_pulses: Tensor,
                                    x0_width: Optional[int]):
        return ops.global_variables(x, graph, ptr, n_pulses, x0_width)

    @staticmethod
    def _linear_chain(seq: torch.nn.Sequential, x: Tensor) -> Tensor:
        mods = list(seq)
        for i in range(0, len(mods), 2):
            w = pad_columns(mods[i].weight, 4)
            if w.shape[1] != x.shape[1]:
                x = pad_columns(x, 4)
            x = ops.linear_act(x, w, mods[i].bias, ops.ACT_RELU)
        return x

    def _post_process(self, skips: List[Tensor], widths: List[int]) -> Tensor:
        """First post-processing Linear as a K-split GEMM over the skip parts."""
        if not _is_linear_relu_chain(self._post_processing):
            x = torch.cat([s[:, :w] for s, w in zip(skips, widths)], dim=1)
            return self._post_processing(x)
        lin = self._post_processing[0]
        blocks, offsets, parts, col, off = [], [], [], 0, 0
        for s, w in zip(skips, widths):
            s = pad_columns(s, 4)
            blk = lin.weight[:, col:col + w]
            if s.shape[1] != w:
                blk = torch.nn.functional.pad(blk, (0, s.shape[1] - w))
            blocks.append(blk)
            offsets.append(off)
            parts.append(s)
            col += w
            off += s.shape[1]
        packed = torch.cat(blocks, dim=1)
        x = ops.multi_linear_act(parts, packed, lin.bias, offsets, ops.ACT_RELU)
        rest = torch.nn.Sequential(*list(self._post_processing)[2:])
        return self._linear_chain(rest, x) if len(rest) else x

    # -- native executor --------------------------------------------------------------------------
    def _executor_config(self):
        """`gnb_dynedge_config` for the fast-path family, or None when this model needs the per-operator route."""
        cached = getattr(self, "_exec_cfg", None)
        if cached is not None:
            return cached if cached != "unsupported" else None
        from graphnet_b200._lib import DynEdgeConfig
        ok = (_is_linear_relu_chain(self._post_processing) and _is_linear_relu_chain(self._readout)
              and all(_hoistable(c.nn) and c.aggr in ("add", "sum") for c in self._conv_layers)
              and len(self._conv_layers) <= 8 and len(self._post_processing) <= 16 and len(self._readout) <= 16
              and 4 <= self._nb_inputs <= 32)
        cols: List[int] = []
        if ok:
            widths = [w for c in self._conv_layers for w in (c.nn[0].out_features, c.nn[2].out_features)]
            widths += [m.out_features for m in self._post_processing if isinstance(m, torch.nn.Linear)]
            widths += [m.out_features for m in self._readout if isinstance(m, torch.nn.Linear)]
            cols = ops.resolve_columns(self._features_subset, min(c.nn[2].out_features for c in self._conv_layers))
            ok = all(w % 4 == 0 for w in widths) and 1 <= len(cols) <= 16
        if not ok:
            object.__setattr__(self, "_exec_cfg", "unsupported")
            return None
        cfg = DynEdgeConfig()
        cfg.nb_inputs, cfg.k, cfg.precision = self._nb_inputs, self._nb_neighbours, 0
        cfg.n_conv = len(self._conv_layers)
        for i, c in enumerate(self._conv_layers):
            cfg.conv_hidden[i], cfg.conv_out[i] = c.nn[0].out_features, c.nn[2].out_features
        posts = [m for m in self._post_processing if isinstance(m, torch.nn.Linear)]
        cfg.n_post = len(posts)
        for i, m in enumerate(posts):
            cfg.post_out[i] = m.out_features
        ros = [m for m in self._readout if isinstance(m, torch.nn.Linear)]
        cfg.n_readout = len(ros)
        for i, m in enumerate(ros):
            cfg.readout_out[i] = m.out_features
        schemes = self._global_pooling_schemes or []
        cfg.n_pool = len(schemes)
        for i, s in enumerate(schemes):
            cfg.pool[i] = ops.POOL[s]
        cfg.globals_after_pooling = int(self._add_global_variables_after_pooling)
        cfg.skip_readout = int(self._skip_readout)
        cfg.n_knn_cols = len(cols)
        for i, c in enumerate(cols):
            cfg.knn_cols[i] = c
        object.__setattr__(self, "_exec_cfg", cfg)
        object.__setattr__(self, "_exec_cols", cols)
        return cfg

    def _executor_params(self) -> List[Tensor]:
        out: List[Tensor] = []
        for c in self._conv_layers:
            out += [c.nn[0].weight, c.nn[0].bias, c.nn[2].weight, c.nn[2].bias]
        for m in self._post_processing:
            if isinstance(m, torch.nn.Linear):
                out += [m.weight, m.bias]
        if not self._skip_readout:
            for m in self._readout:
                if isinstance(m, torch.nn.Linear):
                    out += [m.weight, m.bias]
        return out

    def forward(self, data) -> Tensor:
        """Apply learnable forward pass (dynedge.py:295-349)."""
        x, batch = data.x, data.batch
        if not x.is_cuda:
            raise RuntimeError("graphnet_b200.DynEdge runs on CUDA tensors only (no CPU fallback)")
        n_pulses = data.n_pulses
        nseg = int(n_pulses.numel())
        ptr = getattr(data, "ptr", None)
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, nseg)
        graph = data.knn_graph() if hasattr(data, "knn_graph") else None
        if graph is None:
            graph = ops.KnnGraph.from_edge_index(data.edge_index, x.shape[0], self._nb_neighbours)

        cfg = self._executor_config() if ops.USE_EXECUTOR else None
        if cfg is not None:
            record = {} if getattr(self, "_debug_record", False) else None
            if self._skip_readout:
                out_cols, per_event = self._post_processing_layer_sizes[-1], False
            else:
                out_cols, per_event = self._readout_layer_sizes[-1], bool(self._global_pooling_schemes)
            y = ops.dynedge_execute(cfg, graph, ptr, n_pulses, x, self._exec_cols, out_cols, per_event,
                                    self._executor_params(), record)
            if record is not None:
                self._debug = {"graphs": record["graphs"], "skips": [record["x0"]] + record["ys"],
                               "global_variables": record["g"]}
            return y

        distribute = not self._add_global_variables_after_pooling
        node_width = self._nb_inputs + (self._nb_global_variables if distribute else 0)
        x0_width = ((node_width + 31) // 32) * 32 if distribute else None
        g, x0 = self._calculate_global_variables(x, graph, ptr, n_pulses, x0_width)
        x = x0 if distribute else x.float()

        skips, widths = [x], [node_width]
        nconv = len(self._conv_layers)
        record = getattr(self, "_debug_record", False)
        graphs = [graph]
        for li, conv in enumerate(self._conv_layers):
            x_in = skips[-1]                       # x0 carries zero pad columns (hoisted route accepts them)
            if not _hoistable(conv.nn):
                x_in = x_in[:, :widths[-1]]
            x, graph = conv.forward_table(x_in, graph, ptr, recompute=li + 1 < nconv)
            skips.append(x)
            widths.append(x.shape[1])
            graphs.append(graph)
        if record:   # test hook: the graph fed to every conv, the layer outputs and the global variables
            self._debug = {"graphs": graphs[:nconv], "skips": [s[:, :w] for s, w in zip(skips, widths)],
                           "global_variables": g}

        x = self._post_process(skips, widths)

        if not self._skip_readout:
            if self._global_pooling_schemes:
                x = self._global_pooling(x, ptr=ptr)
                if self._add_global_variables_after_pooling:
                    x = torch.cat([x, g], dim=1)
            if _is_linear_relu_chain(self._readout):
                x = self._linear_chain(self._readout, x)
            else:
                x = self._readout(x)
        return x
