"""The other users of `DynEdgeConv` on the CUDA kernels against the oracle restatements that are pinned
on the reference's own dynedge_jinst.py / particlenet.py / dynedge.py (tests/test_oracle_golden.py): `DynEdgeJINST`
(LeakyReLU MLPs, add), `ParticleNeT` (3-Linear MLPs with BatchNorm1d in training and eval mode, mean aggregation, static
graph with GELU) and DeepIce's DynEdge block (k = 9, GELU, LayerNorm, pulse-level output; icemix.py:100-118).
rel 1e-3 on outputs and every parameter gradient; kNN graphs bit-exact on the kernel's own features."""

import os
from types import SimpleNamespace

import pytest
import torch

from helpers import GOLDEN_DIR, rel_err
from oracle.dynedge_oracle import DynEdgeRef, batch_to_ptr, knn_graph_ref
from oracle.users_oracle import DynEdgeJINSTRef, DynEdgeTITORef, ParticleNeTRef

pytestmark = pytest.mark.gpu
CASES = ["jinst", "particlenet_train", "particlenet_eval", "particlenet_static_gelu", "deepice_dynedge", "tito",
         "tito_residual_no_globals"]


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("case", CASES)
def test_dynedgeconv_users_vs_oracle(built_library, case, precision):
    from graphnet_b200 import Data, ops
    from graphnet_b200.models.gnn import DynEdge, DynEdgeJINST, DynEdgeTITO, ParticleNeT
    from graphnet_b200.models.graphs.edges import KNNEdges
    g = torch.load(os.path.join(GOLDEN_DIR, "users_dynedgeconv.pt"))[case]
    if case == "jinst":
        model, ref, k, cols = DynEdgeJINST(**g["kwargs"]), DynEdgeJINSTRef(**g["kwargs"]), 8, slice(0, 3)
    elif case == "deepice_dynedge":
        model, ref, k, cols = DynEdge(g["nb_inputs"], **g["kwargs"]), DynEdgeRef(g["nb_inputs"], **g["kwargs"]), 9, slice(0, 3)
    elif case.startswith("tito"):       # static graph; EdgeConvTito with max aggregation + per-event transformer (eval mode)
        model, ref, k, cols = DynEdgeTITO(g["nb_inputs"], **g["kwargs"]), DynEdgeTITORef(g["nb_inputs"], **g["kwargs"]), 8, None
    else:
        model, ref = ParticleNeT(g["nb_inputs"], **g["kwargs"]), ParticleNeTRef(g["nb_inputs"], **g["kwargs"])
        k, cols = g["kwargs"]["nb_neighbours"], slice(0, 3)
    assert list(model.state_dict().keys()) == list(ref.state_dict().keys())            # drop-in: same state_dict keys
    model.load_state_dict(g["state_dict"])
    ref = ref.double()
    ref.load_state_dict({k_: (v.double() if v.is_floating_point() else v) for k_, v in g["state_dict"].items()})
    train = case in ("jinst", "particlenet_train", "deepice_dynedge")
    model = model.cuda().train(train)
    ref.train(train)
    model._debug_record = True
    old = ops.PRECISION
    ops.set_precision(precision)
    try:
        x, batch, n_pulses = g["x"], g["batch"], g["n_pulses"]
        data = KNNEdges(k)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
        assert torch.equal(data.edge_index.cpu(), g["edge_index"])                     # the reference run's own initial graph
        y = model(data)
        w = torch.linspace(0.5, 1.5, y.numel()).reshape(y.shape)
        (y * w.cuda()).sum().backward()
        ptr = batch_to_ptr(batch)
        graphs = model._debug["graphs"] if hasattr(model, "_debug") else []
        forced = [None]
        for li in range(1, len(graphs)):
            ei_k = graphs[li].edge_index().cpu()
            if case != "particlenet_static_gelu":                                       # static graph: nothing recomputed
                feats = model._debug["skips"][li].detach().cpu()
                assert torch.equal(ei_k, knn_graph_ref(feats[:, cols], k, ptr=ptr)), f"latent graph {li}"
            forced.append(ei_k)
        d_ref = SimpleNamespace(x=x.double(), edge_index=g["edge_index"], batch=batch, n_pulses=n_pulses)
        y_ref = ref(d_ref) if case.startswith("tito") else ref(d_ref, forced_graphs=forced)
        (y_ref * w.double()).sum().backward()
        err = rel_err(y, y_ref)
        gerr = {}
        for (k_, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                continue
            if float(q.grad.abs().max()) < 1e-9:
                # a Linear bias in front of a training-mode BatchNorm1d: its gradient is identically zero (the batch mean is
                # subtracted); fp32 leaves rounding noise there, which has no relative scale
                assert float(p.grad.abs().max()) < 1e-5, k_
                continue
            gerr[k_] = rel_err(p.grad, q.grad)
        print(f"{case} {precision}: out {err:.2e}, max grad {max(gerr.values()):.2e} ({max(gerr, key=gerr.get)})")
        assert err < 1e-3
        # tito in tf32x3: max aggregation routes its gradient to ONE arg-max edge and the LeakyReLU has a kink -- on 71 pulses
        # one such decision within the forward error (1e-5) of a tie moves a tensor by 1e-2 (measured 6e-4 or 1.0e-2 from
        # build to build); stated there: median 1e-3, every tensor 2e-2. fp32 and every other case: 1e-3 on every tensor.
        med = sorted(gerr.values())[len(gerr) // 2]
        if case == "tito" and precision == "tf32x3":
            assert med < 1e-3 and max(gerr.values()) < 2e-2, gerr
        else:
            assert max(gerr.values()) < 1e-3, gerr
    finally:
        ops.set_precision(old)
