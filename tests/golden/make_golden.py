"""Generate golden vectors by running the REFERENCE's own DynEdge code.

The reference (`/root/reference/src/graphnet/models/gnn/dynedge.py`,
`models/components/layers.py`, `models/utils.py`, `models/gnn/gnn.py`) cannot be
imported as a package in this image: torch_geometric, torch_cluster,
torch_scatter, pytorch_lightning, colorlog, ruamel.yaml ... are not installed
(SURVEY.md section 8c). This script therefore

  1. registers small *shim* modules for exactly the third-party names those four
     files import (`torch_geometric.nn.EdgeConv`, `knn_graph`, `homophily`,
     `torch_scatter.scatter_*`, `pytorch_lightning.LightningModule`, ...),
     written from the published semantics of those packages and deliberately
     NOT sharing code with `oracle/dynedge_oracle.py` where an independent
     formulation is cheap (scatter ops via `Tensor.scatter_reduce`, EdgeConv via
     `index_add_`, homophily via bincount);
  2. registers empty stand-ins for the `graphnet`, `graphnet.models`, ...
     *packages* (so their heavy `__init__`s are not executed) whose `__path__`
     points into `/root/reference/src`, and loads the four reference files
     unmodified from there;
  3. runs the reference `DynEdge` (literal code path, including the dense
     `[N,B]` "distribute" broadcast) on small seeded inputs and stores inputs,
     weights, outputs and parameter gradients under `tests/golden/*.pt`.

What this pins: everything written in the reference's own files. What it does
not pin: the third-party operators themselves (kNN tie order etc.), for which
the reference holds no vectors.

Run (only in the build container, where /root/reference exists):
    python tests/golden/make_golden.py
"""

from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- #
# shims for the absent third-party packages
# --------------------------------------------------------------------------- #
def _knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **_):
    """torch_cluster.knn_graph: brute force, total order (dist, index)."""
    x = x.detach().float()
    n = x.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.int64)
    src, dst = [], []
    for q in range(n):
        members = torch.nonzero(batch == batch[q]).flatten()
        best = []  # sorted list of (dist, idx), strict '>' insertion
        for c in members.tolist():
            d = torch.zeros((), dtype=torch.float32)
            for j in range(x.shape[1]):
                diff = x[c, j] - x[q, j]
                d = d + diff * diff
            d = float(d)
            pos = len(best)
            for e, (bd, _) in enumerate(best):
                if bd > d:
                    pos = e
                    break
            if pos < k + 1:
                best.insert(pos, (d, c))
                best = best[: k + 1]
        for d, c in best:
            if c != q and d < 1e10:
                src.append(c)
                dst.append(q)
    return torch.tensor([src, dst], dtype=torch.int64)


class _MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, x, size=None):
        """PyG MessagePassing.propagate for flow='source_to_target': message(x_i = x[1][dst], x_j = x[0][src]) reduced over
        the edges of every target node; nodes without in-edges get 0 (used by the reference's EdgeConvTito, layers.py:100-106)."""
        src, dst = edge_index[0], edge_index[1]
        msg = self.message(x_i=x[1][dst], x_j=x[0][src])
        n = x[1].shape[0]
        idx = dst.unsqueeze(1).expand_as(msg)
        reduce = {"add": "sum", "sum": "sum", "mean": "mean", "max": "amax", "min": "amin"}[self.aggr]
        return torch.zeros(n, msg.shape[1], dtype=msg.dtype).scatter_reduce(0, idx, msg, reduce, include_self=reduce == "sum")


def _to_dense_batch(x, batch=None):
    """torch_geometric.utils.to_dense_batch: [B, L_max, C] zero padded + bool mask."""
    if batch is None:
        return x.unsqueeze(0), torch.ones(1, x.shape[0], dtype=torch.bool)
    nb = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=nb)
    lmax = int(counts.max())
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(x.shape[0]) - start[batch]
    dense = torch.zeros(nb, lmax, x.shape[1], dtype=x.dtype)
    dense[batch, pos] = x
    mask = torch.zeros(nb, lmax, dtype=torch.bool)
    mask[batch, pos] = True
    return dense, mask


class _EdgeConv(_MessagePassing):
    """PyG EdgeConv: out_i = aggr_j nn(cat[x_i, x_j - x_i])."""

    def __init__(self, nn, aggr="max", **kwargs):
        super().__init__(aggr=aggr, **kwargs)
        self.nn = nn

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        msg = self.nn(torch.cat([x[dst], x[src] - x[dst]], dim=-1))
        n = x.shape[0]
        idx = dst.unsqueeze(1).expand_as(msg)
        if self.aggr in ("add", "sum"):
            return torch.zeros(n, msg.shape[1], dtype=msg.dtype).scatter_reduce(0, idx, msg, "sum", include_self=True)
        if self.aggr == "mean":
            return torch.zeros(n, msg.shape[1], dtype=msg.dtype).scatter_reduce(0, idx, msg, "mean", include_self=False)
        if self.aggr == "max":
            return torch.zeros(n, msg.shape[1], dtype=msg.dtype).scatter_reduce(0, idx, msg, "amax", include_self=False)
        raise ValueError(self.aggr)


def _scatter(reduce):
    def fn(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0
        size = int(index.max()) + 1 if dim_size is None else dim_size
        idx = index.unsqueeze(1).expand_as(src)
        res = torch.zeros(size, src.shape[1], dtype=src.dtype).scatter_reduce(
            0, idx, src, reduce, include_self=False)
        if reduce in ("amin", "amax"):
            return res, torch.zeros(size, src.shape[1], dtype=torch.int64)  # arg unused by DynEdge
        return res
    return fn


def _homophily(edge_index, y, batch=None, method="edge"):
    row, col = edge_index[0], edge_index[1]
    same = (y[row] == y[col]).double()
    nb = int(batch.max()) + 1
    eb = batch[col]
    tot = torch.bincount(eb, weights=same, minlength=nb)
    cnt = torch.bincount(eb, minlength=nb).clamp(min=1)
    return (tot / cnt).float()


class _LightningModule(torch.nn.Module):
    @property
    def device(self):
        return torch.device("cpu")


class _Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Model(torch.nn.Module):
    """Stand-in for graphnet.models.Model (Logger+Configurable+LightningModule)."""

    def __init__(self, *a, **k):
        super().__init__()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims() -> None:
    _mod("torch_geometric")
    _mod("torch_geometric.nn", EdgeConv=_EdgeConv, knn_graph=_knn_graph)
    _mod("torch_geometric.nn.pool", knn_graph=_knn_graph)
    _mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing)
    _mod("torch_geometric.nn.inits", reset=lambda *_: None)
    _mod("torch_geometric.typing", Adj=torch.Tensor, PairTensor=tuple)
    _mod("torch_geometric.data", Data=_Data, Batch=_Data)
    _mod("torch_geometric.utils", homophily=_homophily, to_dense_batch=_to_dense_batch)
    _mod("torch_scatter", scatter_max=_scatter("amax"), scatter_min=_scatter("amin"),
         scatter_sum=_scatter("sum"), scatter_mean=_scatter("mean"))
    _mod("pytorch_lightning", LightningModule=_LightningModule)
    for pkg, sub in [("graphnet", ""), ("graphnet.models", "models"),
                     ("graphnet.models.components", "models/components"),
                     ("graphnet.models.gnn", "models/gnn")]:
        m = _mod(pkg)
        m.__path__ = [os.path.join(REF_SRC, "graphnet", sub)]
    sys.modules["graphnet.models"].Model = _Model


def load_reference_dynedge():
    install_shims()
    return importlib.import_module("graphnet.models.gnn.dynedge").DynEdge


# --------------------------------------------------------------------------- #
# seeded inputs / weights (shared with the tests through this module)
# --------------------------------------------------------------------------- #
def make_events(sizes, nb_inputs, seed):
    """Pulse-map-like events: coarse xyz grid (many exact ties/duplicates)."""
    rng = np.random.default_rng(seed)
    xs = []
    for n in sizes:
        doms = rng.integers(0, 6, size=(max(1, int(np.ceil(0.6 * n))), 3)).astype(np.float32) * 0.25
        pick = rng.integers(0, doms.shape[0], size=n)
        rest = rng.normal(size=(n, nb_inputs - 3)).astype(np.float32)
        xs.append(np.concatenate([doms[pick], rest], axis=1))
    x = torch.from_numpy(np.concatenate(xs, 0))
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    n_pulses = torch.tensor(sizes, dtype=torch.int32)
    return x, batch, n_pulses


def seeded_state_dict(module: torch.nn.Module, seed: int):
    """Weights that do not depend on constructor order: per-key generator."""
    out = {}
    for i, (key, val) in enumerate(sorted(module.state_dict().items())):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if key.endswith("weight") and val.dim() == 2:
            out[key] = (torch.rand(val.shape, generator=g) * 2 - 1) / (val.shape[1] ** 0.5)
        elif key.endswith("weight"):
            out[key] = 1.0 + 0.1 * (torch.rand(val.shape, generator=g) * 2 - 1)
        else:
            out[key] = 0.1 * (torch.rand(val.shape, generator=g) * 2 - 1)
    return out


CASES = {
    # name: (ctor kwargs, event sizes, nb_inputs, store_weights)
    "small_relu_4pool": (dict(dynedge_layer_sizes=[(32, 48), (40, 48)], post_processing_layer_sizes=[40, 32],
                              readout_layer_sizes=[16], global_pooling_schemes=["min", "max", "mean", "sum"]),
                         [2, 5, 9, 10, 40, 33, 12], 7, True),
    "small_gelu_norm_after": (dict(dynedge_layer_sizes=[(24, 32), (24, 32), (16, 32)],
                                   post_processing_layer_sizes=[32], readout_layer_sizes=[24, 8],
                                   global_pooling_schemes=["max", "sum"], add_global_variables_after_pooling=True,
                                   activation_layer="gelu", add_norm_layer=True, nb_neighbours=4),
                              [3, 7, 21, 16], 5, True),
    "small_skip_readout": (dict(dynedge_layer_sizes=[(16, 24)], post_processing_layer_sizes=[24, 16],
                                readout_layer_sizes=[8], global_pooling_schemes=None, skip_readout=True),
                           [4, 11, 9], 4, True),
    "small_node_readout": (dict(dynedge_layer_sizes=[(16, 24), (16, 24)], post_processing_layer_sizes=[16],
                                readout_layer_sizes=[8], global_pooling_schemes=None),
                           [6, 10, 13], 6, True),
    "default_f7": (dict(global_pooling_schemes=["min", "max", "mean", "sum"]), [2, 14, 30, 9, 25], 7, False),
}


def main() -> None:
    DynEdge = load_reference_dynedge()
    for name, (kwargs, sizes, nb_inputs, store_w) in CASES.items():
        torch.manual_seed(0)
        model = DynEdge(nb_inputs, **kwargs).double()   # fp64 run = the target the fp32 paths approximate
        sd = seeded_state_dict(model, seed=7)
        model.load_state_dict({k: v.double() for k, v in sd.items()})
        x, batch, n_pulses = make_events(sizes, nb_inputs, seed=11)
        k = kwargs.get("nb_neighbours", 8)
        edge_index = _knn_graph(x[:, [0, 1, 2]], k, batch)
        out = {}
        for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
            m = model.to(dtype)
            data = _Data(x=x.to(dtype), edge_index=edge_index, batch=batch, n_pulses=n_pulses)
            y = m(data)
            m.zero_grad()
            wsum = torch.linspace(0.5, 1.5, y.numel(), dtype=dtype).reshape(y.shape)
            (y * wsum).sum().backward()
            out[f"out_{tag}"] = y.detach().clone()
            out[f"grads_{tag}"] = {k_: p.grad.detach().clone().float() for k_, p in m.named_parameters()
                                   if p.grad is not None}
        fixture = {"kwargs": kwargs, "sizes": sizes, "nb_inputs": nb_inputs, "x": x, "batch": batch,
                   "n_pulses": n_pulses, "edge_index": edge_index, "weight_seed": 7, **out}
        if store_w:
            fixture["state_dict"] = sd
        else:   # keep the fixture small: only norms of the grads + the outputs
            for tag in ("f64", "f32"):
                fixture[f"grads_{tag}"] = {k_: torch.stack([g.norm(), g.abs().max(), g.flatten()[0]])
                                           for k_, g in out[f"grads_{tag}"].items()}
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(fixture, path)
        print(name, "out", tuple(out["out_f32"].shape), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
