"""CPU oracle for GraphNeT's DynEdge hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference
algorithm. It is the checker for the CUDA path; it is never shipped, never
called from `graphnet_b200/`, and is only imported by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs.

What it restates (paths relative to /root/reference):
  * `src/graphnet/models/gnn/dynedge.py:24-349`      -> `DynEdgeRef`
  * `src/graphnet/models/components/layers.py:20-69` -> `dynedgeconv_ref`
  * `src/graphnet/models/utils.py:13-29`             -> `homophily_ref`, `global_variables_ref`
  * `src/graphnet/models/graphs/edges/edges.py:72-80`-> `knn_graph_ref`

Third-party semantics hard-coded here because the packages are not installed
in this image: torch_cluster `knn_graph` 1.6.x (CUDA
kernel's brute-force scan with strict `>` insertion, i.e. the total order
(distance, index)), PyG `EdgeConv` / `homophily` 2.3+, torch_scatter
`scatter_{min,max,sum,mean}` 2.x.

PINNING STATUS: the model-level structure (global-variable order, the dense
"distribute" broadcast, layer construction, skip-cat order, pooling order,
read-out) is pinned against the reference's *own* `dynedge.py`/`layers.py`/
`utils.py`, executed in this container with shims for the absent third-party
ops (`tests/golden/make_golden.py` -> `tests/golden/*.pt`). At the third-party
operator boundary (kNN tie order, scatter arg choice) parity is UNPINNED: the
reference holds no test vectors for it and the packages cannot be run here.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

KNN_INIT_DIST = 1e10  # torch_cluster knn_cuda.cu initial best distance [EXT]


# --------------------------------------------------------------------------- #
# ptr / batch helpers
# --------------------------------------------------------------------------- #
def batch_to_ptr(batch: Tensor, num_graphs: Optional[int] = None) -> Tensor:
    """`ptr = bucketize(arange(B+1), batch)` as in torch_cluster.knn [EXT]."""
    if num_graphs is None:
        num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
    ar = torch.arange(num_graphs + 1, dtype=batch.dtype, device=batch.device)
    return torch.searchsorted(batch.contiguous(), ar).to(torch.int64)


# --------------------------------------------------------------------------- #
# kNN graph  (layers.py:63-67, edges.py:74-78 + torch_cluster [EXT])
# --------------------------------------------------------------------------- #
def _pair_dist(query: Tensor, cand: Tensor) -> Tensor:
    """L2^2 in fp32, left-to-right over columns, every op rounded (no FMA)."""
    acc = None
    for c in range(query.shape[1]):
        diff = cand[None, :, c] - query[:, None, c]
        sq = diff * diff
        acc = sq if acc is None else acc + sq
    return acc


def knn_graph_ref(feat: Tensor, k: int, batch: Optional[Tensor] = None,
                  ptr: Optional[Tensor] = None, chunk: int = 2048) -> Tensor:
    """`knn_graph(x, k, batch, loop=False, flow='source_to_target')`.

    For each query q of event b: candidates are all nodes of b (q included);
    keep the min(k+1, n_b) smallest under the total order (dist, index); emit
    (neighbour, q) in that order, skipping neighbour == q. Returns int64 [2,E].
    """
    feat = feat.detach().to(torch.float32).contiguous()
    n_total = feat.shape[0]
    if ptr is None:
        if batch is None:
            ptr = torch.tensor([0, n_total], dtype=torch.int64)
        else:
            ptr = batch_to_ptr(batch.cpu())
    ptr_l = ptr.tolist()
    src_parts: List[Tensor] = []
    dst_parts: List[Tensor] = []
    for b in range(len(ptr_l) - 1):
        lo, hi = ptr_l[b], ptr_l[b + 1]
        n = hi - lo
        if n <= 0:
            continue
        cand = feat[lo:hi]
        k1 = min(k + 1, n)
        for q0 in range(0, n, chunk):
            q1 = min(n, q0 + chunk)
            d = _pair_dist(cand[q0:q1], cand)                    # [q, n]
            vals, idx = torch.sort(d, dim=1, stable=True)        # ties: lower index first
            vals, idx = vals[:, :k1], idx[:, :k1]
            qid = torch.arange(q0, q1).unsqueeze(1).expand_as(idx)
            keep = (idx != qid) & (vals < KNN_INIT_DIST)
            src_parts.append(idx[keep] + lo)                     # row-major => ascending distance per q
            dst_parts.append(qid[keep] + lo)
    if not src_parts:
        return torch.zeros(2, 0, dtype=torch.int64)
    return torch.stack([torch.cat(src_parts), torch.cat(dst_parts)]).to(torch.int64)


# --------------------------------------------------------------------------- #
# segment reductions (torch_scatter semantics [EXT])
# --------------------------------------------------------------------------- #
class _SegmentExtreme(torch.autograd.Function):
    """scatter_min / scatter_max: gradient goes to ONE arg (first occurrence)."""

    @staticmethod
    def forward(ctx, x: Tensor, ptr: Tensor, is_max: bool):  # type: ignore
        nseg = ptr.numel() - 1
        out = x.new_zeros(nseg, x.shape[1])
        arg = torch.full((nseg, x.shape[1]), -1, dtype=torch.int64)
        ptr_l = ptr.tolist()
        for b in range(nseg):
            lo, hi = ptr_l[b], ptr_l[b + 1]
            if hi <= lo:
                continue
            seg = x[lo:hi]
            val = seg.max(dim=0).values if is_max else seg.min(dim=0).values
            first = (seg == val.unsqueeze(0)).to(torch.int8).argmax(dim=0)
            out[b] = val
            arg[b] = first + lo
        ctx.save_for_backward(arg)
        ctx.n = x.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, gout: Tensor, _garg):  # type: ignore
        (arg,) = ctx.saved_tensors
        gx = gout.new_zeros(ctx.n, gout.shape[1])
        valid = arg >= 0
        cols = torch.arange(gout.shape[1]).unsqueeze(0).expand_as(arg)
        gx[arg[valid], cols[valid]] = gout[valid]
        return gx, None, None


def segment_pool_ref(x: Tensor, ptr: Tensor, scheme: str) -> Tensor:
    """`scatter_<scheme>(x, batch, dim=0)` for sorted `batch` given as `ptr`."""
    nseg = ptr.numel() - 1
    counts = (ptr[1:] - ptr[:-1])
    if scheme in ("sum", "mean"):
        batch = torch.repeat_interleave(torch.arange(nseg), counts)
        out = x.new_zeros(nseg, x.shape[1]).index_add_(0, batch, x)
        if scheme == "mean":
            out = out / counts.clamp(min=1).to(x.dtype).unsqueeze(1)
        return out
    if scheme == "max":
        return _SegmentExtreme.apply(x, ptr, True)[0]
    if scheme == "min":
        return _SegmentExtreme.apply(x, ptr, False)[0]
    raise ValueError(scheme)


# --------------------------------------------------------------------------- #
# homophily + global variables (models/utils.py:13-29, dynedge.py:266-293)
# --------------------------------------------------------------------------- #
def homophily_ref(edge_index: Tensor, y: Tensor, batch: Tensor, num_graphs: int) -> Tensor:
    """PyG `homophily(edge_index, y, batch, method='edge')` [EXT]."""
    src, dst = edge_index[0], edge_index[1]
    same = (y[src] == y[dst]).to(torch.float32)
    eb = batch[dst]
    total = torch.zeros(num_graphs, dtype=torch.float32).index_add_(0, eb, same)
    count = torch.zeros(num_graphs, dtype=torch.float32).index_add_(0, eb, torch.ones_like(same))
    return total / count.clamp(min=1)


def global_variables_ref(x: Tensor, edge_index: Tensor, batch: Tensor,
                         n_pulses: Tensor, ptr: Tensor) -> Tensor:
    """[mean(x) | h_x h_y h_z h_t | log10(n_pulses)] -> [B, F+5]."""
    nseg = ptr.numel() - 1
    means = segment_pool_ref(x, ptr, "mean")
    hs = [homophily_ref(edge_index, x[:, c].detach(), batch, nseg).reshape(-1, 1).to(x.dtype)
          for c in range(4)]
    # dynedge.py:287 takes log10 of the int32 `n_pulses` tensor itself: torch evaluates that in float32 whatever the dtype of
    # x (visible only in fp64 runs, at 6e-8)
    logn = torch.log10(n_pulses.to(torch.float32) if not n_pulses.is_floating_point() else n_pulses)
    return torch.cat([means] + hs + [logn.to(x.dtype).unsqueeze(1)], dim=1)


# --------------------------------------------------------------------------- #
# EdgeConv / DynEdgeConv (layers.py:55-69 + PyG EdgeConv [EXT])
# --------------------------------------------------------------------------- #
def edgeconv_ref(x: Tensor, edge_index: Tensor, nn: torch.nn.Module, aggr: str) -> Tensor:
    """out_i = AGG_{j in N(i)} nn(cat[x_i, x_j - x_i]); empty neighbourhood -> 0."""
    src, dst = edge_index[0], edge_index[1]
    n = x.shape[0]
    if src.numel() == 0:
        width = nn(torch.zeros(1, 2 * x.shape[1], dtype=x.dtype)).shape[1]
        return x.new_zeros(n, width)
    x_i, x_j = x[dst], x[src]
    msg = nn(torch.cat([x_i, x_j - x_i], dim=-1))
    if aggr in ("add", "sum", "mean"):
        out = msg.new_zeros(n, msg.shape[1]).index_add_(0, dst, msg)
        if aggr == "mean":
            deg = torch.zeros(n, dtype=msg.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=msg.dtype))
            out = out / deg.clamp(min=1).unsqueeze(1)
        return out
    if aggr == "max":
        # edges are grouped by target (knn_graph output); torch_scatter routes the
        # gradient to a single arg-max edge
        order = torch.argsort(dst, stable=True)
        msg_s, dst_s = msg[order], dst[order]
        rowptr = torch.searchsorted(dst_s, torch.arange(n + 1))
        return _SegmentExtreme.apply(msg_s, rowptr, True)[0]
    raise ValueError(aggr)


def dynedgeconv_ref(x: Tensor, edge_index: Tensor, nn: torch.nn.Module, aggr: str, k: int,
                    features_subset: Union[slice, Sequence[int]], batch: Optional[Tensor],
                    ptr: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    out = edgeconv_ref(x, edge_index, nn, aggr)
    new_ei = knn_graph_ref(out[:, features_subset], k, batch=batch, ptr=ptr)
    return out, new_ei


# --------------------------------------------------------------------------- #
# DynEdge (dynedge.py) with identical ctor and state_dict keys
# --------------------------------------------------------------------------- #
class _ConvRef(torch.nn.Module):
    def __init__(self, nn: torch.nn.Module, aggr: str, nb_neighbors: int, features_subset):
        super().__init__()
        self.nn = nn
        self.aggr = aggr
        self.nb_neighbors = nb_neighbors
        self.features_subset = features_subset


class DynEdgeRef(torch.nn.Module):
    """Restatement of `DynEdge` (dynedge.py:21-349)."""

    def __init__(self, nb_inputs: int, *, nb_neighbours: int = 8,
                 features_subset=None, dynedge_layer_sizes=None,
                 post_processing_layer_sizes=None, readout_layer_sizes=None,
                 global_pooling_schemes=None, add_global_variables_after_pooling: bool = False,
                 activation_layer: Optional[str] = None, add_norm_layer: bool = False,
                 skip_readout: bool = False, literal_distribute: bool = False):
        super().__init__()
        if features_subset is None:
            features_subset = slice(0, 3)                                   # dynedge.py:78-79
        if dynedge_layer_sizes is None:
            dynedge_layer_sizes = [(128, 256), (336, 256), (336, 256), (336, 256)]
        if post_processing_layer_sizes is None:
            post_processing_layer_sizes = [336, 256]
        if readout_layer_sizes is None:
            readout_layer_sizes = [128]
        if isinstance(global_pooling_schemes, str):
            global_pooling_schemes = [global_pooling_schemes]
        if global_pooling_schemes is not None:
            for s in global_pooling_schemes:
                assert s in ("min", "max", "sum", "mean")
        if add_global_variables_after_pooling:
            assert global_pooling_schemes
        if activation_layer is None or activation_layer.lower() == "relu":
            act: torch.nn.Module = torch.nn.ReLU()
        elif activation_layer.lower() == "gelu":
            act = torch.nn.GELU()
        else:
            raise ValueError(f"Activation layer {activation_layer} not supported.")
        self._act = act
        self._nb_inputs = nb_inputs
        self._nb_global = 5 + nb_inputs
        self._k = nb_neighbours
        self._subset = features_subset
        self._schemes = global_pooling_schemes
        self._after = add_global_variables_after_pooling
        self._skip_readout = skip_readout
        self._literal = literal_distribute

        nb_in_feat = nb_inputs + (0 if self._after else self._nb_global)   # dynedge.py:186-188
        self._conv_layers = torch.nn.ModuleList()
        latent = nb_in_feat
        for sizes in dynedge_layer_sizes:                                   # dynedge.py:192-213
            layers: List[torch.nn.Module] = []
            dims = [latent] + list(sizes)
            for ix, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
                if ix == 0:
                    a *= 2
                layers.append(torch.nn.Linear(a, b))
                if add_norm_layer:
                    layers.append(torch.nn.LayerNorm(b))
                layers.append(act)
            self._conv_layers.append(_ConvRef(torch.nn.Sequential(*layers), "add",
                                              nb_neighbours, features_subset))
            latent = dims[-1]
        latent = sum(s[-1] for s in dynedge_layer_sizes) + nb_in_feat      # dynedge.py:216-219
        layers = []
        dims = [latent] + list(post_processing_layer_sizes)
        for a, b in zip(dims[:-1], dims[1:]):
            layers.append(torch.nn.Linear(a, b))
            if add_norm_layer:
                layers.append(torch.nn.LayerNorm(b))
            layers.append(act)
        self._post_processing = torch.nn.Sequential(*layers)
        npool = len(global_pooling_schemes) if global_pooling_schemes else 1
        latent = dims[-1] * npool + (self._nb_global if self._after else 0)  # dynedge.py:234-241
        layers = []
        dims = [latent] + list(readout_layer_sizes)
        for a, b in zip(dims[:-1], dims[1:]):
            layers.append(torch.nn.Linear(a, b))
            layers.append(act)
        self._readout = torch.nn.Sequential(*layers)

    @staticmethod
    def _apply_forcing_last_relu(seq: torch.nn.Sequential, x: Tensor, mask: Optional[Tensor]):
        """seq(x), with the ReLU decisions of the LAST activation taken from `mask` when given: y = z * mask."""
        if mask is None or len(seq) == 0 or not isinstance(seq[-1], torch.nn.ReLU):
            return seq(x), None
        z = seq[:-1](x)
        return z * mask.to(z.dtype), z

    def forward(self, data, forced_graphs: Optional[List[Tensor]] = None,
                return_intermediates: bool = False, forced_output_mask: Optional[Tensor] = None):
        """dynedge.py:295-349. `forced_graphs[l]` (optional) replaces the graph
        used as INPUT of conv layer l (l>=1) -- used to compare against a kernel
        run whose latent features (and thus kNN graphs) differ by rounding.
        `forced_output_mask` (optional, bool, shape of the output) replaces the ReLU decisions of the model's LAST
        activation (read-out, or post-processing with skip_readout): the gradient of a ReLU network is discontinuous where a
        pre-activation crosses zero, and with B x 128 read-out units a single unit within rounding of zero moves a whole
        bias-gradient entry by 1 / B -- the same teacher-forcing as for the graphs; callers assert that the forced decisions
        differ from the oracle's own only within rounding of zero (`final_pre` in the intermediates)."""
        x, edge_index, batch = data.x, data.edge_index, data.batch
        nseg = int(batch.max().item()) + 1
        ptr = batch_to_ptr(batch, nseg)
        g = global_variables_ref(x, edge_index, batch, data.n_pulses, ptr)   # :300-305
        if not self._after:
            if self._literal:                                                # :308-317 (dense form)
                distribute = (batch.unsqueeze(1) == torch.unique(batch).unsqueeze(0)).type(x.dtype)
                gd = torch.sum(distribute.unsqueeze(2) * g.unsqueeze(0), dim=1)
            else:
                gd = g[batch]
            x = torch.cat((x, gd), dim=1)                                    # :319
        skips = [x]
        graphs = [edge_index]
        for li, conv in enumerate(self._conv_layers):                       # :322-325
            if forced_graphs is not None and li < len(forced_graphs) and forced_graphs[li] is not None:
                edge_index = forced_graphs[li]
            x = edgeconv_ref(x, edge_index, conv.nn, conv.aggr)
            if li + 1 < len(self._conv_layers) or return_intermediates:
                edge_index = knn_graph_ref(x[:, self._subset], self._k, ptr=ptr)
            graphs.append(edge_index)
            skips.append(x)
        x = torch.cat(skips, dim=1)                                          # :328
        final_pre = None
        if self._skip_readout:
            x, final_pre = self._apply_forcing_last_relu(self._post_processing, x, forced_output_mask)   # :331
        else:
            x = self._post_processing(x)                                     # :331
        post = x
        if not self._skip_readout:
            if self._schemes:
                x = torch.cat([segment_pool_ref(x, ptr, s) for s in self._schemes], dim=1)  # :251-264
                if self._after:
                    x = torch.cat([x, g], dim=1)                             # :337-344
            x, final_pre = self._apply_forcing_last_relu(self._readout, x, forced_output_mask)            # :347
        if return_intermediates:
            return x, {"global_variables": g, "skips": skips, "graphs": graphs, "post": post, "final_pre": final_pre}
        return x


# --------------------------------------------------------------------------------------------- #
# graph definition in front of the path -- checker for models/graphs/device.py
# --------------------------------------------------------------------------------------------- #
def standardize_icecube86_ref(raw: torch.Tensor, feature_names) -> torch.Tensor:
    """`Detector._standardize` (reference detector/detector.py:63-77) with the IceCube86 map (detector/icecube.py:21-48):
    one callable per named column, fp32, an unknown column is a KeyError. Pinned bit-for-bit on
    tests/golden/detector_icecube86.pt (the reference's own two files run here)."""
    fmap = {
        "dom_x": lambda x: x / 500.0, "dom_y": lambda x: x / 500.0, "dom_z": lambda x: x / 500.0,       # icecube.py:35-36
        "dom_time": lambda x: (x - 1.0e04) / 3.0e4,                                                       # :38-39
        "charge": lambda x: torch.log10(x),                                                               # :41-42
        "rde": lambda x: (x - 1.25) / 0.25,                                                               # :44-45
        "pmt_area": lambda x: x / 0.05,                                                                   # :47-48
        "hlc": lambda x: x,                                                                               # detector.py:79-81
    }
    out = raw.clone()
    for idx, name in enumerate(feature_names):
        out[:, idx] = fmap[name](raw[:, idx])
    return out


def collate_ref(n_pulses: torch.Tensor):
    """`batch` and `ptr` of `Batch.from_data_list` (reference data/dataloader.py:12-18) for events of the given sizes."""
    sizes = n_pulses.to(torch.int64)
    ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
    batch = torch.repeat_interleave(torch.arange(sizes.numel(), dtype=torch.int64), sizes)
    return batch, ptr
