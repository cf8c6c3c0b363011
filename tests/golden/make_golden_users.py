"""Golden vectors for the other users of `DynEdgeConv`, produced by the REFERENCE's own, unmodified
`src/graphnet/models/gnn/dynedge_jinst.py`, `src/graphnet/models/gnn/particlenet.py` and -- for DeepIce's DynEdge block --
`dynedge.py` with the arguments of `src/graphnet/models/gnn/icemix.py:100-118`, loaded under the shims of make_golden.py.
fp64 runs, outputs + parameter gradients; ParticleNeT in training mode (batch statistics, dropout 0) and in eval mode.

Run (only in the build container, where /root/reference exists):  python tests/golden/make_golden_users.py
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def run(model, x, batch, n_pulses, k, train=True):
    model = model.double()
    sd = mg.seeded_state_dict(model, seed=7)
    for key in list(sd):                       # BatchNorm bookkeeping keeps its defaults
        if "running_" in key or "num_batches" in key:
            sd[key] = model.state_dict()[key]
    model.load_state_dict({k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()})
    model.train(train)
    edge_index = mg._knn_graph(x[:, [0, 1, 2]], k, batch)
    data = mg._Data(x=x.double(), edge_index=edge_index, batch=batch, n_pulses=n_pulses)
    y = model(data)
    model.zero_grad()
    w = torch.linspace(0.5, 1.5, y.numel(), dtype=torch.float64).reshape(y.shape)
    (y * w).sum().backward()
    grads = {k_: p.grad.detach().clone().float() for k_, p in model.named_parameters() if p.grad is not None}
    return {"x": x, "batch": batch, "n_pulses": n_pulses, "edge_index": edge_index, "state_dict": sd,
            "out_f64": y.detach().clone(), "grads_f64": grads}


def main() -> None:
    mg.install_shims()
    jinst = importlib.import_module("graphnet.models.gnn.dynedge_jinst").DynEdgeJINST
    pnet = importlib.import_module("graphnet.models.gnn.particlenet").ParticleNeT
    dynedge = importlib.import_module("graphnet.models.gnn.dynedge").DynEdge
    out = {}
    x, batch, n_pulses = mg.make_events([3, 12, 30, 9, 17], 7, seed=21)
    torch.manual_seed(0)
    out["jinst"] = dict(run(jinst(7, layer_size_scale=1), x, batch, n_pulses, 8), kwargs=dict(nb_inputs=7, layer_size_scale=1))
    kw = dict(nb_neighbours=4, dynedge_layer_sizes=[(16, 16, 16), (24, 24, 24)], readout_layer_sizes=[16],
              global_pooling_schemes=["mean", "max"], dropout_readout=0.0)
    x, batch, n_pulses = mg.make_events([6, 14, 25, 11], 5, seed=22)
    torch.manual_seed(0)
    out["particlenet_train"] = dict(run(pnet(5, **kw), x, batch, n_pulses, 4, train=True), kwargs=kw, nb_inputs=5)
    torch.manual_seed(0)
    out["particlenet_eval"] = dict(run(pnet(5, **kw), x, batch, n_pulses, 4, train=False), kwargs=kw, nb_inputs=5)
    kw2 = dict(nb_neighbours=4, dynamic=False, dynedge_layer_sizes=[(16, 16), (24, 24)], readout_layer_sizes=[8],
               global_pooling_schemes=None, add_batchnorm_layer=False, activation_layer="gelu", skip_readout=True)
    torch.manual_seed(0)
    out["particlenet_static_gelu"] = dict(run(pnet(5, **kw2), x, batch, n_pulses, 4, train=False), kwargs=kw2, nb_inputs=5)
    # DeepIce's DynEdge block (icemix.py:100-118) at reduced widths: k = 9, GELU, LayerNorm, pulse-level output
    kw3 = dict(nb_neighbours=9, post_processing_layer_sizes=[40, 24], dynedge_layer_sizes=[(16, 32), (40, 32), (40, 32), (40, 32)],
               global_pooling_schemes=None, activation_layer="gelu", add_norm_layer=True, skip_readout=True)
    x, batch, n_pulses = mg.make_events([4, 13, 28, 10], 9, seed=23)
    torch.manual_seed(0)
    out["deepice_dynedge"] = dict(run(dynedge(9, **kw3), x, batch, n_pulses, 9), kwargs=kw3, nb_inputs=9)
    # DynEdgeTITO (dynedge_kaggle_tito.py + the reference's own EdgeConvTito / DynTrans in layers.py) at reduced widths, eval
    # mode (the TransformerEncoderLayer's dropout is random in training mode)
    tito = importlib.import_module("graphnet.models.gnn.dynedge_kaggle_tito").DynEdgeTITO
    kw4 = dict(dyntrans_layer_sizes=[(32, 32), (32, 32), (48, 48)], global_pooling_schemes=["max", "mean"], n_head=4,
               post_processing_layer_sizes=[40, 24], readout_layer_sizes=[24, 8])
    x, batch, n_pulses = mg.make_events([3, 12, 30, 9, 17], 6, seed=24)
    torch.manual_seed(0)
    out["tito"] = dict(run(tito(6, **kw4), x, batch, n_pulses, 8, train=False), kwargs=kw4, nb_inputs=6)
    kw5 = dict(dyntrans_layer_sizes=[(6, 6)], global_pooling_schemes=["max"], n_head=2, use_global_features=False,
               use_post_processing_layers=False, readout_layer_sizes=[8])
    torch.manual_seed(0)
    out["tito_residual_no_globals"] = dict(run(tito(6, **kw5), x, batch, n_pulses, 8, train=False), kwargs=kw5, nb_inputs=6)
    torch.save(out, os.path.join(HERE, "users_dynedgeconv.pt"))
    for k_, v in out.items():
        print(k_, tuple(v["out_f64"].shape), float(v["out_f64"].abs().max()))


if __name__ == "__main__":
    main()
