"""BASELINE configs[0]: the reference's own example data -- `data/examples/sqlite/prometheus/prometheus-events.db`, 50 events,
1 872 pulses (19 % duplicate-xyz pulses, 8 events with fewer than 9 pulses: the k + 1 duplicate quirk and the short-event
cases of the kNN contract), F = 4, batch 16 (examples/04_training/01_train_dynedge.py:85,113-142,195,223) -- through
KNNGraph(Prometheus) -> collate -> device KNNEdges -> DynEdge(4) forward + backward against the oracle, in every precision mode.
The pulse table travels as tests/golden/prometheus_events.npz (written by tests/golden/make_prometheus_fixture.py from the
reference's data base and its own, unmodified detector/prometheus.py)."""

import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, namespace, oracle_on_kernel_decisions, rel_err
from oracle.dynedge_oracle import DynEdgeRef, batch_to_ptr, knn_graph_ref

pytestmark = pytest.mark.gpu
# Batches of 16 events / ~600 pulses (and one of 2 events / 61 pulses): fp32 meets rel 1e-3; tf32x3 (single-pass tf32 backward
# GEMMs) is stated at 2e-3 on batches this small (measured 1.4e-3; 7.5e-4 on the 512-event training batch). The single-pass
# tf32 mode is not run here: its forward rounding flips ReLU decisions all over a 600-pulse network (measured 2e-2).
GRAD_TOL = {"fp32": 1e-3, "tf32x3": 2e-3, "mixed16": 2e-3}


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "mixed16"])
def test_prometheus_example_epoch_vs_oracle(built_library, precision):
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN_DIR[:-len("/golden")]))
    import bench_workloads
    from graphnet_b200 import ops
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.tasks import EnergyReconstruction
    fx = np.load(os.path.join(GOLDEN_DIR, "prometheus_events.npz"))
    batches, definition = bench_workloads.prometheus_batches(16)
    assert [int(b.n_pulses.numel()) for b in batches] == [16, 16, 16, 2] and definition.nb_outputs == 4
    # the host-side detector reproduces the reference's own standardisation bit for bit
    assert torch.equal(torch.cat([b.x for b in batches]), torch.from_numpy(fx["standardized"]))
    kwargs = dict(global_pooling_schemes=["min", "max", "mean", "sum"])
    torch.manual_seed(0)
    ref = DynEdgeRef(4, **kwargs)
    head_ref = EnergyReconstruction(128)
    model = DynEdge(4, **kwargs)
    model.load_state_dict(ref.state_dict())
    model = model.cuda()
    head = EnergyReconstruction(128)
    head.load_state_dict(head_ref.state_dict())
    head = head.cuda()
    model._debug_record = True
    ref, head_ref = ref.double(), head_ref.double()
    old = ops.PRECISION
    ops.set_precision(precision)
    try:
        worst_out, worst_grad, forced_total = 0.0, 0.0, 0
        for hb in batches:
            assert hb.edge_index is None                                    # edges deferred on the CPU (dataloader workers)
            dev = definition.build_edges(hb.to("cuda"))
            x, batch, n_pulses = hb.x, hb.batch, hb.n_pulses
            ptr = batch_to_ptr(batch)
            ei0 = knn_graph_ref(x[:, :3], 8, ptr=ptr)
            assert torch.equal(dev.edge_index.cpu(), ei0)                   # initial graph: bit-exact (duplicates, short events)
            model.zero_grad(set_to_none=True)
            head.zero_grad(set_to_none=True)
            feat = model(dev)
            pred = head(feat)
            loss = head.compute_loss(pred, hb.total_energy.float().cuda())
            loss.backward()
            forced = [None]
            for li in range(1, 4):                                          # all four graphs torch.equal
                feats = model._debug["skips"][li].detach().cpu()
                ei_k = model._debug["graphs"][li].edge_index().cpu()
                assert torch.equal(ei_k, knn_graph_ref(feats[:, :3], 8, ptr=ptr)), f"latent graph {li}"
                forced.append(ei_k)
            for p in list(ref.parameters()) + list(head_ref.parameters()):
                p.grad = None
            feat_ref, _, nforced = oracle_on_kernel_decisions(ref, namespace(x=x.double(), edge_index=ei0, batch=batch,
                                                                             n_pulses=n_pulses), forced, feat, precision)
            forced_total += nforced
            pred_ref = head_ref(feat_ref)
            loss_ref = head_ref.compute_loss(pred_ref, hb.total_energy.double())
            loss_ref.backward()
            worst_out = max(worst_out, rel_err(pred, pred_ref), rel_err(loss, loss_ref))
            gerr = {k: rel_err(p.grad, q.grad) for (k, p), (_, q) in zip(list(model.named_parameters()) + list(head.named_parameters()),
                                                                        list(ref.named_parameters()) + list(head_ref.named_parameters()))}
            worst_grad = max(worst_grad, max(gerr.values()))
            wk = max(gerr, key=gerr.get)
            print(f"  batch of {int(n_pulses.numel())} events / {x.shape[0]} pulses: worst grad {wk} {gerr[wk]:.2e}; "
                  + ", ".join(f"{k_.split('.', 1)[-1]}={v:.1e}" for k_, v in gerr.items() if v > 1e-3))
        print(f"prometheus50 {precision}: predictions / loss {worst_out:.2e}, max grad {worst_grad:.2e}, "
              f"read-out ReLU decisions taken from the kernel: {forced_total}")
        assert worst_out < 1e-3
        assert worst_grad < GRAD_TOL[precision]
    finally:
        ops.set_precision(old)
