"""`DynEdgeConv`: EdgeConv followed by a kNN recompute in the layer's output space.

Reference: src/graphnet/models/components/layers.py:20-69 (a PyG `EdgeConv`
whose `forward` also calls `knn_graph` on `x[:, features_subset]`).

B200 design. The per-edge MLP input `[x_i, x_j - x_i]` is never materialised
when the MLP has the DynEdge shape `Linear, ReLU, Linear, ReLU`: the first
Linear is hoisted to nodes (`W1 [x_i; x_j-x_i] + b1 = (W1a-W1b) x_i + b1 + W1b x_j`),
so the only per-edge dense contraction left is the second Linear, and the
k-neighbour aggregation reads each message once. Any other `nn` (LayerNorm,
GELU, BatchNorm, 3-layer MLPs ...) runs through the generic route: gather
kernel -> the user's `nn` -> aggregation kernel. Both routes are CUDA only.
"""

from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from graphnet_b200 import ops
from graphnet_b200.models.model import Model


def _hoistable(nn: Any) -> bool:
    if not isinstance(nn, torch.nn.Sequential) or len(nn) != 4:
        return False
    return (isinstance(nn[0], torch.nn.Linear) and isinstance(nn[1], torch.nn.ReLU)
            and isinstance(nn[2], torch.nn.Linear) and isinstance(nn[3], torch.nn.ReLU))


def _has_batch_statistics(nn: Any) -> bool:
    return isinstance(nn, torch.nn.Module) and any(
        isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in nn.modules())


def pad_columns(t: Tensor, mult: int) -> Tensor:
    pad = (-t.shape[1]) % mult
    return t if pad == 0 else torch.nn.functional.pad(t, (0, pad))


class DynEdgeConv(Model):
    """Dynamical edge convolution layer."""

    def __init__(
        self,
        nn: Callable,
        aggr: str = "max",
        nb_neighbors: int = 8,
        features_subset: Optional[Union[Sequence[int], slice]] = None,
        **kwargs: Any,
    ):
        if features_subset is None:
            features_subset = slice(None)
        assert isinstance(features_subset, (list, slice))
        assert aggr in ops.AGGR, f"aggr={aggr!r} not supported"
        super().__init__()
        self.nn = nn
        self.aggr = aggr
        self.nb_neighbors = nb_neighbors
        self.features_subset = features_subset

    # -- EdgeConv: out_i = AGG_j nn([x_i, x_j - x_i]) --------------------------------------------
    def edge_conv(self, x: Tensor, graph: ops.KnnGraph) -> Tensor:
        if _hoistable(self.nn):
            lin1, lin2 = self.nn[0], self.nn[2]
            c = lin1.in_features // 2
            cp = x.shape[1]                       # x may carry zero pad columns beyond c
            w1 = lin1.weight
            wa, wb = w1[:, :c], w1[:, c:]
            wcat = torch.cat([wa - wb, wb], dim=0)                       # [2H, c]
            if cp != c:
                wcat = torch.nn.functional.pad(wcat, (0, cp - c))
            wcat = pad_columns(wcat, 4)
            if wcat.shape[1] != cp:
                x = pad_columns(x, 4)
            bcat = None
            if lin1.bias is not None:
                bcat = torch.cat([lin1.bias, torch.zeros_like(lin1.bias)])
            pq = ops.linear_act(x, wcat, bcat, ops.ACT_NONE, round_out=False)   # [N, 2H] = [P | Q], added in fp32
            if lin1.out_features % 4 == 0 and lin2.out_features % 4 == 0 and self.aggr in ("add", "sum", "mean"):
                return ops.edgeconv_hoisted(pq, lin2.weight, lin2.bias, graph, self.aggr)
            if lin1.out_features % 4 == 0:
                h = ops.edge_hidden(pq, graph, ops.ACT_RELU)             # [N*W, H]
                m = ops.linear_act(h, lin2.weight, lin2.bias, ops.ACT_RELU)
                return ops.edge_aggregate(m, graph, self.aggr)
        # generic route: any callable `nn`
        u = ops.edge_cat(x, graph)
        if _has_batch_statistics(self.nn):
            valid = (torch.arange(graph.width, device=x.device).unsqueeze(0) < graph.deg.unsqueeze(1)).flatten()
            mv = self.nn(u[valid])
            m = mv.new_zeros(u.shape[0], mv.shape[1])
            m[valid] = mv
        else:
            m = self.nn(u)
        return ops.edge_aggregate(m.float(), graph, self.aggr)

    def recompute_graph(self, x: Tensor, ptr: Tensor) -> ops.KnnGraph:
        cols = ops.resolve_columns(self.features_subset, x.shape[1])
        return ops.knn_table(x, cols, ptr, self.nb_neighbors)

    def forward_table(self, x: Tensor, graph: ops.KnnGraph, ptr: Tensor,
                      recompute: bool = True) -> Tuple[Tensor, Optional[ops.KnnGraph]]:
        x = self.edge_conv(x, graph)
        return x, (self.recompute_graph(x, ptr) if recompute else None)

    def forward(self, x: Tensor, edge_index, batch: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Reference signature: returns the new features and the recomputed `edge_index`."""
        n = x.shape[0]
        graph = edge_index if isinstance(edge_index, ops.KnnGraph) else ops.KnnGraph.from_edge_index(edge_index, n)
        if batch is None:
            ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
        else:
            ptr = ops.batch_to_ptr(batch, int(batch.max().item()) + 1)
        x, new_graph = self.forward_table(x, graph, ptr)
        return x, new_graph.edge_index()


def _act_code(m: torch.nn.Module):
    """ops.ACT_* of an activation module the kernels implement (None: something else)."""
    if isinstance(m, torch.nn.ReLU):
        return ops.ACT_RELU
    if isinstance(m, torch.nn.LeakyReLU) and m.negative_slope == 0.01:
        return ops.ACT_LEAKY
    if isinstance(m, torch.nn.Identity):
        return ops.ACT_NONE
    return None


class EdgeConvTito(Model):
    """EdgeConv of the TITO solution: out_i = AGG_j nn([x_i, x_j - x_i, x_j]) (reference: layers.py:72-114, a PyG
    `MessagePassing` with `aggr="max"` by default).

    B200 design: when `nn` starts with a Linear over the 3C message (it always does in `DynTrans`), that Linear is hoisted
    from edges to nodes exactly like DynEdge's -- W1 [x_i; x_j - x_i; x_j] + b1 = (W1a - W1b) x_i + b1 + (W1b + W1c) x_j --
    so the [E, 3C] message is never built: one [N, C] x [C, 2H] GEMM, the gather-add kernel (`edge_hidden`, no activation),
    the rest of `nn` on the [E, H] rows and the arg-routed max-aggregation kernel. Other `nn`s take the generic route."""

    def __init__(self, nn: Callable, aggr: str = "max", **kwargs: Any):
        assert aggr in ops.AGGR, f"aggr={aggr!r} not supported"
        super().__init__()
        self.nn = nn
        self.aggr = aggr

    def edge_conv(self, x: Tensor, graph: ops.KnnGraph) -> Tensor:
        c = x.shape[1]
        seq = self.nn if isinstance(self.nn, torch.nn.Sequential) else None
        first = seq[0] if seq is not None and len(seq) and isinstance(seq[0], torch.nn.Linear) else None
        if first is not None and first.in_features == 3 * c and first.out_features % 4 == 0:
            w = first.weight
            wa, wb, wc = w[:, :c], w[:, c:2 * c], w[:, 2 * c:]
            wcat = pad_columns(torch.cat([wa - wb, wb + wc], dim=0), 4)                  # [2H, c]
            xp = pad_columns(x, 4)
            bcat = None if first.bias is None else torch.cat([first.bias, torch.zeros_like(first.bias)])
            pq = ops.linear_act(xp, wcat, bcat, ops.ACT_NONE, round_out=False)           # [N, 2H] = [P | Q]
            acts = [_act_code(m_) for m_ in list(seq)[1:]]
            if (self.aggr == "max" and len(seq) == 4 and isinstance(seq[2], torch.nn.Linear) and acts[0] is not None
                    and acts[2] is not None and ops.edgeconv_hoisted_max_ok(graph, first.out_features, seq[2].out_features)):
                # Linear, act, Linear, act with max aggregation (every DynTrans layer of DynEdgeTITO): the second Linear, its
                # activation and the maximum over the k slots run as ONE tcgen05 kernel, the backward is arg-routed
                return ops.edgeconv_hoisted_max(pq, seq[2].weight, seq[2].bias, graph, acts[0], acts[2])
            a1 = ops.edge_hidden(pq, graph, ops.ACT_NONE)                                # [N*W, H]: W1 msg + b1 per edge slot
            m = a1
            for layer in list(seq)[1:]:
                m = layer(m)
        else:
            u = ops.edge_cat(x, graph)                                                   # [x_i | x_j - x_i]
            m = self.nn(torch.cat([u, u[:, c:] + u[:, :c]], dim=1))
        return ops.edge_aggregate(m.float(), graph, self.aggr)

    def forward(self, x: Tensor, edge_index) -> Tensor:
        graph = edge_index if isinstance(edge_index, ops.KnnGraph) else ops.KnnGraph.from_edge_index(edge_index, x.shape[0])
        return self.edge_conv(x, graph)


def to_dense_events(x: Tensor, ptr: Tensor) -> Tuple[Tensor, Tensor]:
    """`torch_geometric.utils.to_dense_batch` for sorted events given as `ptr`: ([B, L_max, C] zero padded, mask [B, L_max])."""
    sizes = ptr[1:] - ptr[:-1]
    nseg, lmax = int(sizes.numel()), int(sizes.max().item()) if sizes.numel() else 0
    batch = torch.repeat_interleave(torch.arange(nseg, device=x.device), sizes)
    pos = torch.arange(x.shape[0], device=x.device) - ptr[:-1][batch]
    dense = x.new_zeros(nseg, lmax, x.shape[1])
    dense[batch, pos] = x
    mask = torch.zeros(nseg, lmax, dtype=torch.bool, device=x.device)
    mask[batch, pos] = True
    return dense, mask


# The reference applies `TransformerEncoder` to a zero-padded dense batch [events, longest event, C] (layers.py:190-195). Every
# part of an encoder layer except the attention itself is per token, so here the four Linear layers run on the packed token
# list [N, C] through the tensor-core Linear kernels (no padding rows: 2.2 x fewer rows at <= 256 pulses per event; fp32-grade
# tcgen05 instead of torch's fp32 SIMT GEMMs) and only q / k / v are padded for the per-event attention. False, or precision
# mode "fp32" (the reference-arithmetic mode), = the literal module call on the padded batch.
TRANSFORMER_ON_TOKENS = True


def _encoder_layer_supported(layer: torch.nn.Module) -> bool:
    return (isinstance(layer, torch.nn.TransformerEncoderLayer) and not layer.norm_first
            and layer.activation in (torch.nn.functional.relu,) and layer.self_attn.in_proj_weight is not None
            and layer.self_attn.batch_first and layer.self_attn.bias_k is None and not layer.self_attn.add_zero_attn)


def _linear_tokens(x: Tensor, lin_w: Tensor, lin_b: Optional[Tensor], act: int) -> Tensor:
    """act(x W^T + b) on the Linear kernels; wide layers in chunks of 1024 outputs (the split-operand GEMM's limit)."""
    if lin_w.shape[0] <= 1024:
        return ops.linear_act(x, lin_w, lin_b, act, round_out=False)
    outs = [ops.linear_act(x, lin_w[o:o + 1024], None if lin_b is None else lin_b[o:o + 1024], act, round_out=False)
            for o in range(0, lin_w.shape[0], 1024)]
    return torch.cat(outs, dim=1)


def encoder_layer_on_tokens(layer: torch.nn.TransformerEncoderLayer, x: Tensor, ptr: Tensor) -> Tensor:
    """A post-norm `TransformerEncoderLayer` (ReLU feed-forward) applied per event to packed tokens x [N, C]: same parameters and
    arithmetic as `layer(dense, src_key_padding_mask=~mask)[mask]`."""
    attn = layer.self_attn
    n, c = x.shape
    heads = attn.num_heads
    qkv = _linear_tokens(x, attn.in_proj_weight, attn.in_proj_bias, ops.ACT_NONE)                 # [N, 3C]
    dense, mask = to_dense_events(qkv, ptr)                                                      # [B, L, 3C]
    b, l = mask.shape
    q, k, v = dense.view(b, l, 3, heads, c // heads).permute(2, 0, 3, 1, 4)                      # [B, heads, L, C / heads] each
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask[:, None, None, :],
                                                         dropout_p=attn.dropout if layer.training else 0.0)
    o = o.permute(0, 2, 1, 3).reshape(b, l, c)[mask]                                             # back to tokens
    o = _linear_tokens(o, attn.out_proj.weight, attn.out_proj.bias, ops.ACT_NONE)
    x = layer.norm1(x + layer.dropout1(o))
    ff = _linear_tokens(x, layer.linear1.weight, layer.linear1.bias, ops.ACT_RELU)
    ff = _linear_tokens(layer.dropout(ff), layer.linear2.weight, layer.linear2.bias, ops.ACT_NONE)
    return layer.norm2(x + layer.dropout2(ff))


class DynTrans(EdgeConvTito):
    """`dynTrans1` layer of the TITO solution (reference: layers.py:117-197): EdgeConvTito with a LeakyReLU MLP, residual
    connection when the widths agree, LayerNorm, and one TransformerEncoder layer applied per event (padded dense batch with
    a key-padding mask, like the reference's `to_dense_batch`). Same attribute names (`nn`, `norm1`,
    `_transformer_encoder`) and therefore `state_dict` keys. The graph is static: this fork never recomputes kNN here."""

    def __init__(self, layer_sizes: Optional[List[int]] = None, aggr: str = "max",
                 features_subset: Optional[Union[Sequence[int], slice]] = None, n_head: int = 8, **kwargs: Any):
        if features_subset is None:
            features_subset = slice(None)
        assert isinstance(features_subset, (list, slice))
        if layer_sizes is None:
            layer_sizes = [256, 256, 256]
        layers: List[torch.nn.Module] = []
        for ix, (nb_in, nb_out) in enumerate(zip(layer_sizes[:-1], layer_sizes[1:])):
            layers.append(torch.nn.Linear(3 * nb_in if ix == 0 else nb_in, nb_out))
            layers.append(torch.nn.LeakyReLU())
        d_model = layer_sizes[-1]
        super().__init__(nn=torch.nn.Sequential(*layers), aggr=aggr, **kwargs)
        self.features_subset = features_subset
        self.norm1 = torch.nn.LayerNorm(d_model, eps=1e-5)
        encoder_layer = torch.nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, batch_first=True, norm_first=False)
        self._transformer_encoder = torch.nn.TransformerEncoder(encoder_layer, num_layers=1)

    def forward_table(self, x: Tensor, graph: ops.KnnGraph, ptr: Tensor) -> Tensor:
        x_out = self.edge_conv(x, graph)
        x = x + x_out if x_out.shape[-1] == x.shape[-1] else x_out
        x = self.norm1(x)
        enc = self._transformer_encoder
        if (TRANSFORMER_ON_TOKENS and ops.PRECISION != "fp32" and enc.norm is None
                and all(_encoder_layer_supported(layer) for layer in enc.layers)):
            for layer in enc.layers:
                x = encoder_layer_on_tokens(layer, x, ptr)
            return x
        dense, mask = to_dense_events(x, ptr)
        dense = enc(dense, src_key_padding_mask=~mask)
        return dense[mask]

    def forward(self, x: Tensor, edge_index, batch: Optional[Tensor] = None) -> Tensor:
        n = x.shape[0]
        graph = edge_index if isinstance(edge_index, ops.KnnGraph) else ops.KnnGraph.from_edge_index(edge_index, n)
        if batch is None:
            ptr = torch.tensor([0, n], dtype=torch.int64, device=x.device)
        else:
            ptr = ops.batch_to_ptr(batch, int(batch.max().item()) + 1)
        return self.forward_table(x, graph, ptr)
