"""CPU oracle for the other users of `DynEdgeConv` -- TEST INFRASTRUCTURE ONLY (never imported by the
product; see oracle/dynedge_oracle.py for the rules).

Restates, on top of the operator semantics of oracle/dynedge_oracle.py:
  * `DynEdgeJINST.forward`  (/root/reference/src/graphnet/models/gnn/dynedge_jinst.py:103-152)
  * `ParticleNeT.forward`   (/root/reference/src/graphnet/models/gnn/particlenet.py:179-255)
PINNING STATUS: both are pinned on golden vectors produced by the reference's own, unmodified files under the shims of
tests/golden/make_golden.py (tests/golden/make_golden_users.py -> users_*.pt, checked by tests/test_oracle_golden.py); the
third-party operator boundary (kNN tie order, scatter semantics) stays unpinned as stated in dynedge_oracle.py.
"""

from __future__ import annotations

from typing import List, Optional

import torch

from oracle.dynedge_oracle import (batch_to_ptr, edgeconv_ref, homophily_ref, knn_graph_ref, segment_pool_ref)


class DynEdgeJINSTRef(torch.nn.Module):
    def __init__(self, nb_inputs: int, layer_size_scale: int = 4):
        super().__init__()
        c = layer_size_scale
        l1, l2, l3, l4, l5, l6 = nb_inputs, c * 16 * 2, c * 32 * 2, c * 42 * 2, c * 32 * 2, c * 16 * 2     # :31-39

        class _Conv(torch.nn.Module):
            def __init__(self, nn):
                super().__init__()
                self.nn = nn

        def block(n_in, hidden):
            return _Conv(torch.nn.Sequential(torch.nn.Linear(n_in * 2, hidden), torch.nn.LeakyReLU(),
                                             torch.nn.Linear(hidden, l3), torch.nn.LeakyReLU()))
        self.conv_add1, self.conv_add2 = block(l1, l2), block(l3, l4)                                       # :48-94
        self.conv_add3, self.conv_add4 = block(l3, l4), block(l3, l4)
        self.nn1 = torch.nn.Linear(l3 * 4 + l1, l4)                                                         # :97-100
        self.nn2 = torch.nn.Linear(l4, l5)
        self.nn3 = torch.nn.Linear(4 * l5 + 5, l6)
        self.lrelu = torch.nn.LeakyReLU()

    def forward(self, data, forced_graphs: Optional[List] = None, return_graphs: bool = False):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        nseg = int(batch.max().item()) + 1
        ptr = batch_to_ptr(batch, nseg)
        h = [homophily_ref(edge_index, x[:, c].detach(), batch, nseg).to(x.dtype) for c in range(4)]       # :108 (x, y, z, t)
        outs, graphs, cur = [], [edge_index], x
        for li, conv in enumerate((self.conv_add1, self.conv_add2, self.conv_add3, self.conv_add4)):      # :110-113
            if forced_graphs is not None and li < len(forced_graphs) and forced_graphs[li] is not None:
                edge_index = forced_graphs[li]
            cur = edgeconv_ref(cur, edge_index, conv.nn, "add")
            if li < 3 or return_graphs:
                edge_index = knn_graph_ref(cur[:, 0:3], 8, ptr=ptr)
            graphs.append(edge_index)
            outs.append(cur)
        x = torch.cat([x] + outs, dim=1)                                                                    # :116
        x = self.nn2(self.lrelu(self.nn1(x)))                                                               # :119-121
        pooled = [segment_pool_ref(x, ptr, s) for s in ("max", "min", "sum", "mean")]                       # :124-127
        x = torch.cat(pooled + [h[3].reshape(-1, 1), h[0].reshape(-1, 1), h[1].reshape(-1, 1), h[2].reshape(-1, 1),
                                data.n_pulses.reshape(-1, 1).to(x.dtype)], dim=1)                           # :130-143
        x = self.lrelu(self.nn3(self.lrelu(x)))                                                             # :146-150
        return (x, graphs) if return_graphs else x


class ParticleNeTRef(torch.nn.Module):
    def __init__(self, nb_inputs: int, *, nb_neighbours: int = 16, features_subset=None, dynamic: bool = True,
                 dynedge_layer_sizes=None, readout_layer_sizes=None, global_pooling_schemes="mean",
                 activation_layer: Optional[str] = "relu", add_batchnorm_layer: bool = True, dropout_readout: float = 0.1,
                 skip_readout: bool = False):
        super().__init__()
        if features_subset is None:
            features_subset = slice(0, 3)
        if dynedge_layer_sizes is None:
            dynedge_layer_sizes = [(64, 64, 64), (128, 128, 128), (256, 256, 256)]
        if readout_layer_sizes is None:
            readout_layer_sizes = [256]
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        act = torch.nn.ReLU() if activation_layer is None or activation_layer.lower() == "relu" else torch.nn.GELU()
        self._k, self._subset, self._dynamic = nb_neighbours, features_subset, dynamic
        self._schemes, self._skip_readout = global_pooling_schemes, skip_readout

        class _Conv(torch.nn.Module):
            def __init__(self, nn):
                super().__init__()
                self.nn = nn
        self._conv_layers = torch.nn.ModuleList()
        width = nb_inputs
        for sizes in dynedge_layer_sizes:                                                                   # :186-207
            layers, dims = [], [width] + list(sizes)
            for ix, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
                layers.append(torch.nn.Linear(2 * a if ix == 0 else a, b))
                if add_batchnorm_layer:
                    layers.append(torch.nn.BatchNorm1d(b))
                layers.append(act)
            self._conv_layers.append(_Conv(torch.nn.Sequential(*layers)))
            width = dims[-1]
        npool = len(global_pooling_schemes) if global_pooling_schemes else 1                               # :209-216
        dims = [width * npool] + list(readout_layer_sizes)
        ro = []
        for a, b in zip(dims[:-1], dims[1:]):
            ro += [torch.nn.Linear(a, b), act, torch.nn.Dropout(dropout_readout)]
        self._readout = torch.nn.Sequential(*ro)

    def forward(self, data, forced_graphs: Optional[List] = None, return_graphs: bool = False):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        ptr = batch_to_ptr(batch, int(batch.max().item()) + 1)
        graphs = [edge_index]
        for li, conv in enumerate(self._conv_layers):                                                       # :240-244
            if forced_graphs is not None and li < len(forced_graphs) and forced_graphs[li] is not None:
                edge_index = forced_graphs[li]
            x = edgeconv_ref(x, edge_index, conv.nn, "mean")
            if self._dynamic and (li + 1 < len(self._conv_layers) or return_graphs):
                edge_index = knn_graph_ref(x[:, self._subset], self._k, ptr=ptr)
            graphs.append(edge_index)
        if not self._skip_readout:                                                                          # :246-253
            if self._schemes:
                x = torch.cat([segment_pool_ref(x, ptr, s) for s in self._schemes], dim=1)
            x = self._readout(x)
        return (x, graphs) if return_graphs else x


# --------------------------------------------------------------------------------------------------------------------- #
# TITO family: EdgeConvTito / DynTrans (/root/reference/src/graphnet/models/components/layers.py:72-197)
# and DynEdgeTITO (/root/reference/src/graphnet/models/gnn/dynedge_kaggle_tito.py:31-278). Pinned like the classes above.
# --------------------------------------------------------------------------------------------------------------------- #
def edgeconv_tito_ref(x, edge_index, nn, aggr: str = "max"):
    """out_i = AGG_j nn([x_i, x_j - x_i, x_j]) (layers.py:100-112); empty neighbourhood -> 0; max routes its gradient to one
    arg-max edge (first occurrence in the target-grouped edge list) [EXT torch_scatter / PyG]."""
    from oracle.dynedge_oracle import _SegmentExtreme
    src, dst = edge_index[0], edge_index[1]
    n = x.shape[0]
    x_i, x_j = x[dst], x[src]
    msg = nn(torch.cat([x_i, x_j - x_i, x_j], dim=-1))
    if aggr in ("add", "sum"):
        return msg.new_zeros(n, msg.shape[1]).index_add_(0, dst, msg)
    order = torch.argsort(dst, stable=True)
    rowptr = torch.searchsorted(dst[order].contiguous(), torch.arange(n + 1))
    return _SegmentExtreme.apply(msg[order], rowptr, aggr == "max")[0]


class DynTransRef(torch.nn.Module):
    def __init__(self, layer_sizes, aggr: str = "max", n_head: int = 8):
        super().__init__()
        layers = []
        for ix, (a, b) in enumerate(zip(layer_sizes[:-1], layer_sizes[1:])):                                # layers.py:151-159
            layers += [torch.nn.Linear(3 * a if ix == 0 else a, b), torch.nn.LeakyReLU()]
        self.nn = torch.nn.Sequential(*layers)
        self.aggr = aggr
        d_model = layer_sizes[-1]
        self.norm1 = torch.nn.LayerNorm(d_model, eps=1e-5)                                                  # :167
        enc = torch.nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, batch_first=True, norm_first=False)   # :170-175
        self._transformer_encoder = torch.nn.TransformerEncoder(enc, num_layers=1)

    def forward(self, x, edge_index, batch):
        x_out = edgeconv_tito_ref(x, edge_index, self.nn, self.aggr)                                        # :183
        x = x + x_out if x_out.shape[-1] == x.shape[-1] else x_out                                          # :185-188
        x = self.norm1(x)                                                                                   # :190
        # to_dense_batch + key-padding mask (:193-195), event by event: padding never influences a valid position
        counts = torch.bincount(batch)
        outs, lo = [], 0
        for c in counts.tolist():
            outs.append(self._transformer_encoder(x[lo:lo + c].unsqueeze(0)).squeeze(0))
            lo += c
        return torch.cat(outs, dim=0)


class DynEdgeTITORef(torch.nn.Module):
    def __init__(self, nb_inputs: int, features_subset=None, dyntrans_layer_sizes=None, global_pooling_schemes=("max",),
                 use_global_features: bool = True, use_post_processing_layers: bool = True, post_processing_layer_sizes=None,
                 readout_layer_sizes=None, n_head: int = 8, nb_neighbours: int = 8):
        super().__init__()
        if dyntrans_layer_sizes is None:
            dyntrans_layer_sizes = [(256, 256)] * 4
        if post_processing_layer_sizes is None:
            post_processing_layer_sizes = [336, 256]
        if readout_layer_sizes is None:
            readout_layer_sizes = [256, 128]
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        self._schemes = list(global_pooling_schemes)
        self._use_globals, self._use_post = use_global_features, use_post_processing_layers
        act = torch.nn.LeakyReLU()
        self._conv_layers = torch.nn.ModuleList()
        width = nb_inputs
        for sizes in dyntrans_layer_sizes:                                                                  # tito.py:161-170
            self._conv_layers.append(DynTransRef([width] + list(sizes), "max", n_head))
            width = sizes[-1]
        if use_post_processing_layers:                                                                      # :172-186
            dims, layers = [width] + list(post_processing_layer_sizes), []
            for a, b in zip(dims[:-1], dims[1:]):
                layers += [torch.nn.Linear(a, b), act]
            self._post_processing = torch.nn.Sequential(*layers)
            width = dims[-1]
        width = width * len(self._schemes) + ((5 + nb_inputs) if use_global_features else 0)                 # :190-198
        dims, layers = [width] + list(readout_layer_sizes), []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [torch.nn.Linear(a, b), act]
        self._readout = torch.nn.Sequential(*layers)

    def forward(self, data):
        from oracle.dynedge_oracle import global_variables_ref
        x, edge_index, batch = data.x, data.edge_index, data.batch
        nseg = int(batch.max().item()) + 1
        ptr = batch_to_ptr(batch, nseg)
        g = global_variables_ref(x, edge_index, batch, data.n_pulses, ptr) if self._use_globals else None   # :229-262
        for conv in self._conv_layers:                                                                      # :264-265 (static graph)
            x = conv(x, edge_index, batch)
        if self._use_post:
            x = self._post_processing(x)
        x = torch.cat([segment_pool_ref(x, ptr, s) for s in self._schemes], dim=1)                          # :270
        if self._use_globals:
            x = torch.cat([x, g], dim=1)
        return self._readout(x)
