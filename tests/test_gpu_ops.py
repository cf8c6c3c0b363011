"""Parity of the individual CUDA operators (global variables, EdgeConv pieces, pooling, dense layers)
against the oracle / plain fp32 torch on the same seeded inputs. Tolerance for floating point:
rel 1e-5 (fp32 SIMT mode: same products, different summation order); integers bit-exact."""

import numpy as np
import pytest
import torch

from helpers import rel_err, tie_heavy_events
from oracle.dynedge_oracle import (batch_to_ptr, edgeconv_ref, global_variables_ref, knn_graph_ref,
                                   segment_pool_ref)

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _graph(x, ptr, k):
    from graphnet_b200 import ops
    return ops.knn_table(x.cuda(), [0, 1, 2], ptr.cuda(), k)


def test_batch_to_ptr(built_library):
    from graphnet_b200 import ops
    batch = torch.tensor([0, 0, 2, 2, 2, 5])
    assert ops.batch_to_ptr(batch.cuda(), 7).cpu().tolist() == [0, 2, 2, 5, 5, 5, 6, 6]
    assert ops.batch_to_ptr(batch.cuda(), 7).cpu().tolist() == batch_to_ptr(batch, 7).tolist()


@pytest.mark.parametrize("nf", [4, 7, 16])
def test_global_variables_and_distribute(built_library, nf):
    from graphnet_b200 import ops
    sizes = [6, 15, 2, 300, 1, 41]
    x, batch, n_pulses = tie_heavy_events(sizes, nf, seed=nf)
    ptr = batch_to_ptr(batch)
    ei = knn_graph_ref(x[:, :3], 8, ptr=ptr)
    g_ref = global_variables_ref(x, ei, batch, n_pulses, ptr)
    graph = _graph(x, ptr, 8)
    width = ((2 * nf + 5 + 31) // 32) * 32
    g, x0 = ops.global_variables(x.cuda(), graph, ptr.cuda(), n_pulses.cuda(), width)
    assert torch.equal(g[:, nf:nf + 4].cpu(), g_ref[:, nf:nf + 4])             # homophily: exact
    assert rel_err(g, g_ref) < TOL
    x0_ref = torch.cat([x, g.cpu()[batch]], dim=1)
    assert torch.equal(x0[:, :2 * nf + 5].cpu(), x0_ref) and float(x0[:, 2 * nf + 5:].abs().max()) == 0.0


@pytest.mark.parametrize("scheme_list", [["min", "max", "mean", "sum"], ["sum"], ["max", "min"]])
def test_segment_pool_forward_backward(built_library, scheme_list):
    from graphnet_b200 import ops
    rng = np.random.default_rng(0)
    ptr = torch.tensor([0, 10, 10, 35, 60, 61, 700])
    x = torch.from_numpy(rng.integers(-3, 4, size=(700, 100)).astype(np.float32)) + \
        torch.from_numpy(rng.normal(size=(700, 100)).astype(np.float32)) * (rng.random((700, 100)) > 0.5)
    xr = x.clone().requires_grad_(True)
    ref = torch.cat([segment_pool_ref(xr, ptr, s) for s in scheme_list], dim=1)
    w = torch.from_numpy(rng.normal(size=tuple(ref.shape)).astype(np.float32))
    (ref * w).sum().backward()
    xg = x.cuda().requires_grad_(True)
    out = ops.segment_pool(xg, ptr.cuda(), scheme_list)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < TOL
    assert rel_err(xg.grad, xr.grad) < TOL


@pytest.mark.parametrize("aggr", ["add", "mean", "max"])
def test_edgeconv_hoisted_and_generic_match_literal_oracle(built_library, aggr):
    from graphnet_b200.models.components.layers import DynEdgeConv
    sizes = [1, 2, 5, 9, 10, 64, 130, 12]
    x, batch, _ = tie_heavy_events(sizes, 20, seed=4)
    x[-12:] = x[-12]                                             # duplicate quirk: k+1 edges
    ptr = batch_to_ptr(batch)
    ei = knn_graph_ref(x[:, :3], 8, ptr=ptr)
    torch.manual_seed(1)
    for hidden, act in [((40, 24), torch.nn.ReLU()), ((40, 24), torch.nn.GELU())]:   # hoisted / generic
        nn = torch.nn.Sequential(torch.nn.Linear(40, hidden[0]), act, torch.nn.Linear(hidden[0], hidden[1]), act)
        xr = x.clone().requires_grad_(True)
        ref = edgeconv_ref(xr, ei, nn, aggr)
        w = torch.linspace(-1, 1, ref.numel()).reshape(ref.shape)
        nn.zero_grad()
        (ref * w).sum().backward()
        gref = {k: p.grad.clone() for k, p in nn.named_parameters()}
        import copy
        conv = DynEdgeConv(copy.deepcopy(nn), aggr=aggr, nb_neighbors=8, features_subset=slice(0, 3)).cuda()
        xg = x.cuda().requires_grad_(True)
        graph = _graph(x, ptr, 8)
        out, new_graph = conv.forward_table(xg, graph, ptr.cuda())
        conv.zero_grad()
        (out * w.cuda()).sum().backward()
        assert rel_err(out, ref) < TOL
        assert rel_err(xg.grad, xr.grad) < 5 * TOL
        for k, p in conv.nn.named_parameters():
            assert rel_err(p.grad, gref[k]) < 5 * TOL, k
        # recomputed graph == oracle kNN on the kernel's own output features (bit-exact)
        assert torch.equal(new_graph.edge_index().cpu(), knn_graph_ref(out.detach().cpu()[:, :3], 8, ptr=ptr))
    # public reference signature: (x, edge_index, batch) -> (x, edge_index)
    out2, ei2 = conv(x.cuda(), ei.cuda(), batch.cuda())
    assert ei2.dtype == torch.int64 and ei2.shape[0] == 2 and rel_err(out2, out) < 1e-6


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (77, 19, 38), (513, 336, 1043), (4096, 256, 336), (130, 128, 20)])
def test_linear_forward_backward_vs_torch(built_library, m, n, k):
    from graphnet_b200 import ops
    torch.manual_seed(m)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    for act in (ops.ACT_NONE, ops.ACT_RELU):
        xr, wr, br = (t.clone().double().requires_grad_(True) for t in (x, w, b))
        ref = torch.nn.functional.linear(xr, wr, br)
        if act:
            ref = torch.relu(ref)
        g = torch.randn(m, n)
        (ref * g.double()).sum().backward()
        xg, wg, bg = (t.cuda().requires_grad_(True) for t in (x, w, b))
        out = ops.linear_act(xg, wg, bg, act)
        (out * g.cuda()).sum().backward()
        assert rel_err(out, ref) < TOL
        assert rel_err(xg.grad, xr.grad) < TOL and rel_err(wg.grad, wr.grad) < TOL and rel_err(bg.grad, br.grad) < TOL


def test_multi_linear_is_linear_on_concatenation(built_library):
    from graphnet_b200 import ops
    torch.manual_seed(0)
    parts = [torch.randn(300, w) for w in (32, 256, 256)]
    w = torch.randn(336, 544) / 23.0
    b = torch.randn(336)
    ref = torch.relu(torch.cat(parts, 1).double() @ w.double().t() + b.double())
    out = ops.multi_linear_act([p.cuda() for p in parts], w.cuda(), b.cuda(), [0, 32, 288], ops.ACT_RELU)
    assert rel_err(out, ref) < TOL


def test_fused_task_heads_match_torch_heads(built_library):
    """csrc/task_heads.cu against the plain-torch heads / losses of graphnet_b200.tasks evaluated in fp64 (those are
    pinned on the reference's own loss functions by tests/test_tasks.py): loss, predictions and all gradients."""
    from graphnet_b200.tasks import DirectionReconstructionWithKappa, EnergyReconstruction, FusedEnergyDirectionTask
    torch.manual_seed(3)
    nev, hdim = 517, 128
    fused = FusedEnergyDirectionTask(hdim).cuda()
    h = (torch.randn(nev, hdim) * torch.logspace(-2, 0.5, nev).unsqueeze(1)).cuda().requires_grad_(True)
    energy = (10 ** (torch.rand(nev) * 4)).cuda()
    direction = torch.nn.functional.normalize(torch.randn(nev, 3), dim=1).cuda()
    loss, pe, pd = fused(h, energy, direction)
    loss.backward()
    # fp64 reference with the same weights
    e64, d64 = EnergyReconstruction(hdim).double(), DirectionReconstructionWithKappa(hdim).double()
    e64.load_state_dict({k: v.double().cpu() for k, v in fused.energy.state_dict().items()})
    d64.load_state_dict({k: v.double().cpu() for k, v in fused.direction.state_dict().items()})
    h64 = h.detach().double().cpu().requires_grad_(True)
    pe64, pd64 = e64(h64), d64(h64)
    loss64 = e64.compute_loss(pe64, energy.double().cpu()) + d64.compute_loss(pd64, direction.double().cpu())
    loss64.backward()
    assert abs(float(loss) - float(loss64)) < 1e-5 * max(1.0, abs(float(loss64)))
    assert rel_err(pe, pe64) < 1e-5 and rel_err(pd, pd64) < 1e-5
    assert rel_err(h.grad, h64.grad) < 2e-4
    for pf, pr in zip(list(fused.energy.parameters()) + list(fused.direction.parameters()),
                      list(e64.parameters()) + list(d64.parameters())):
        assert rel_err(pf.grad, pr.grad) < 2e-4, pf.shape


@pytest.mark.gpu
@pytest.mark.parametrize("eps,wd", [(1e-3, 0.0), (1e-8, 0.01)])
def test_flat_adam_matches_torch_adam(built_library, eps, wd):
    """FlatAdam (one launch over the flat parameter / gradient buffers) follows torch.optim.Adam, the optimizer the
    reference configures (easy_model.py:215-219; lr = 1e-3, eps = 1e-3 in examples/04_training/01_train_dynedge.py)."""
    import copy
    from graphnet_b200.distributed import FlatAdam, FlatGradAllReduce
    torch.manual_seed(0)
    mod = torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.ReLU(), torch.nn.Linear(13, 5)).cuda()   # 174 values: tail of 2
    ref = copy.deepcopy(mod)
    red = FlatGradAllReduce(mod.parameters())
    opt = FlatAdam(red, lr=1e-3, eps=eps, weight_decay=wd)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3, eps=eps, weight_decay=wd)
    for p, q in zip(mod.parameters(), ref.parameters()):
        assert torch.equal(p, q) and p.data_ptr() >= opt.flat_p.data_ptr()      # parameters alias the flat buffer
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        gflat = torch.randn(red.flat.numel(), device="cuda", generator=g) * (10.0 ** (step - 3))   # identical gradients
        red.flat.copy_(gflat)
        off = 0
        for q in ref.parameters():
            q.grad = gflat[off:off + q.numel()].view_as(q).clone()
            off += q.numel()
        opt.step(zero_grad=True)
        ref_opt.step()
        assert float(red.flat.abs().max()) == 0.0                              # zeroed behind the read
        for p, q in zip(mod.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), (step, float((p - q).abs().max()))
