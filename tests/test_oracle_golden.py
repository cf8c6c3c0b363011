"""The oracle against the golden vectors produced by the reference's own DynEdge code
(tests/golden/make_golden.py runs /root/reference/src/graphnet/models/gnn/dynedge.py with shims for the
absent third-party operators). CPU only."""

import os

import pytest
import torch

from helpers import GOLDEN_DIR, golden_files, load_golden, namespace, rel_err, seeded_state_dict
from oracle.dynedge_oracle import DynEdgeRef, knn_graph_ref


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-3])
@pytest.mark.parametrize("literal", [False, True], ids=["gather", "dense_distribute"])
def test_oracle_matches_reference_golden(path, literal):
    fx = load_golden(path)
    model = DynEdgeRef(fx["nb_inputs"], literal_distribute=literal, **fx["kwargs"])
    model.load_state_dict(fx.get("state_dict") or seeded_state_dict(model, fx["weight_seed"]))
    k = fx["kwargs"].get("nb_neighbours", 8)
    edge_index = knn_graph_ref(fx["x"][:, [0, 1, 2]], k, batch=fx["batch"])
    assert torch.equal(edge_index, fx["edge_index"])          # integer work: bit-exact
    data = namespace(x=fx["x"], edge_index=edge_index, batch=fx["batch"], n_pulses=fx["n_pulses"])
    y = model(data)
    w = torch.linspace(0.5, 1.5, y.numel()).reshape(y.shape)
    (y * w).sum().backward()
    assert rel_err(y, fx["out_f32"]) < 1e-6                   # same fp32 op order as the reference
    assert rel_err(y, fx["out_f64"]) < 1e-5
    for key, p in model.named_parameters():
        if p.grad is None:
            assert key not in fx["grads_f32"]
            continue
        g = fx["grads_f32"][key]
        if g.shape == p.grad.shape:
            assert rel_err(p.grad, g) < 1e-5, key
        else:   # compact fixture: [norm, max|.|, first element]
            mine = torch.stack([p.grad.norm(), p.grad.abs().max()])
            assert rel_err(mine, g[:2]) < 1e-5, key


def test_state_dict_keys_match_reference_layout():
    model = DynEdgeRef(7, global_pooling_schemes=["min", "max", "mean", "sum"])
    keys = list(model.state_dict().keys())
    assert "_conv_layers.0.nn.0.weight" in keys and "_conv_layers.3.nn.2.bias" in keys
    assert "_post_processing.0.weight" in keys and "_post_processing.2.bias" in keys and "_readout.0.weight" in keys
    sd = model.state_dict()
    assert tuple(sd["_conv_layers.0.nn.0.weight"].shape) == (128, 38)
    assert tuple(sd["_conv_layers.1.nn.0.weight"].shape) == (336, 512)
    assert tuple(sd["_post_processing.0.weight"].shape) == (336, 1043)
    assert tuple(sd["_readout.0.weight"].shape) == (128, 1024)
    assert sum(p.numel() for p in model.parameters()) == 1382192


def test_detector_standardisation_pinned_on_reference_icecube86():
    """Oracle restatement, the host-side `IceCube86` mirror and its (kind, subtract, divide) table all reproduce the
    reference's own detector/icecube.py + detector.py:_standardize output bit for bit (tests/golden/make_golden_detector.py)."""
    import numpy as np
    from graphnet_b200.models.detector import IceCube86
    from oracle.dynedge_oracle import standardize_icecube86_ref
    gold = torch.load(os.path.join(GOLDEN_DIR, "detector_icecube86.pt"))
    raw, want, names = gold["raw"], gold["standardized"], gold["features"]
    assert torch.equal(standardize_icecube86_ref(raw, names), want)
    det = IceCube86()
    assert torch.equal(det(raw, names), want)
    kinds, subs, divs = det.standardisation_table(names)
    got = raw.clone()
    for c, (k, a, b) in enumerate(zip(kinds, subs, divs)):
        if k == 1:
            got[:, c] = (raw[:, c] - np.float32(a)) / np.float32(b)
        elif k == 2:
            got[:, c] = torch.log10(raw[:, c])
    assert torch.equal(got, want)
    with pytest.raises(KeyError):                  # detector.py:70-76: no silent pass-through
        det(raw, names[:-1] + ["not_a_feature"])
    with pytest.raises(KeyError):
        det.standardisation_table(["not_a_feature"])


def test_percentile_clusters_pinned_on_reference_utils():
    """`PercentileClusters` (host-side node definition of BASELINE config #4) reproduces the reference's own
    `cluster_summarize_with_percentiles` bit for bit in float64 (tests/golden/make_golden_nodes.py), keeps its output
    feature names, and flows through `KNNGraph` (float32 nodes, n_pulses = raw pulse count, edges deferred on the CPU)."""
    from graphnet_b200.models.detector import IdentityDetector
    from graphnet_b200.models.graphs import KNNGraph
    from graphnet_b200.models.graphs.nodes import PercentileClusters
    gold = torch.load(os.path.join(GOLDEN_DIR, "nodes_percentile_clusters.pt"))
    names, cluster_on, pcts = gold["features"], gold["cluster_on"], gold["percentiles"]
    for case in gold["cases"]:
        node_def = PercentileClusters(cluster_on=cluster_on, percentiles=pcts, add_counts=case["add_counts"],
                                      input_feature_names=names)
        data, out_names = node_def(case["x"])
        assert data.x.dtype == torch.float64 and torch.equal(data.x, case["nodes"])
        assert len(out_names) == case["nodes"].shape[1] == 3 + 4 * 3 + int(case["add_counts"])
        assert out_names[:4] == ["dom_x", "dom_y", "dom_z", "dom_time_pct10"] and (out_names[-1] == "counts") == case["add_counts"]
    case = gold["cases"][-2]
    definition = KNNGraph(detector=IdentityDetector(), node_definition=PercentileClusters(cluster_on, pcts),
                          input_feature_names=names)
    graph = definition(case["x"].numpy(), names)
    assert definition.nb_outputs == 16 and graph.x.dtype == torch.float32
    assert torch.equal(graph.x, case["nodes"].float()) and int(graph.n_pulses) == case["x"].shape[0]
    assert torch.equal(graph["counts"], graph.x[:, -1])


def _users_gold():
    return torch.load(os.path.join(GOLDEN_DIR, "users_dynedgeconv.pt"))


@pytest.mark.parametrize("case", ["jinst", "particlenet_train", "particlenet_eval", "particlenet_static_gelu", "deepice_dynedge",
                                  "tito", "tito_residual_no_globals"])
def test_users_oracle_reproduces_the_reference_models(case):
    """oracle/users_oracle.py (DynEdgeJINST, ParticleNeT) and DynEdgeRef with DeepIce's arguments against golden vectors from
    the reference's own dynedge_jinst.py / particlenet.py / dynedge.py (tests/golden/make_golden_users.py): fp64, 1e-10."""
    from types import SimpleNamespace
    from oracle.dynedge_oracle import DynEdgeRef
    from oracle.users_oracle import DynEdgeJINSTRef, DynEdgeTITORef, ParticleNeTRef
    g = _users_gold()[case]
    if case == "jinst":
        model = DynEdgeJINSTRef(**g["kwargs"])
    elif case.startswith("tito"):
        model = DynEdgeTITORef(g["nb_inputs"], **g["kwargs"])
    elif case == "deepice_dynedge":
        model = DynEdgeRef(g["nb_inputs"], **g["kwargs"])
    else:
        model = ParticleNeTRef(g["nb_inputs"], **g["kwargs"])
    model = model.double()
    model.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in g["state_dict"].items()})
    model.train(case in ("jinst", "particlenet_train", "deepice_dynedge"))
    data = SimpleNamespace(x=g["x"].double(), edge_index=g["edge_index"], batch=g["batch"], n_pulses=g["n_pulses"])
    y = model(data)
    w = torch.linspace(0.5, 1.5, y.numel(), dtype=torch.float64).reshape(y.shape)
    (y * w).sum().backward()
    # TITO: the reference pads events into one dense batch (masked attention), the oracle runs event by event: 1e-9
    assert torch.allclose(y, g["out_f64"], rtol=1e-9 if case.startswith("tito") else 1e-10, atol=1e-12)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grads_f64"])
    for k, v in g["grads_f64"].items():
        assert torch.allclose(grads[k].float(), v, rtol=2e-5, atol=1e-7), k
