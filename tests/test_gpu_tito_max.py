"""Max aggregation fused into the tcgen05 GEMM epilogue + arg-routed backward (north_star bullet 2; the per-edge half of
`EdgeConvTito`, reference models/components/layers.py:72-114 with aggr="max" as instantiated at
models/gnn/dynedge_kaggle_tito.py:157-162).

* bit-exact on integer-valued operands (products and sums exact in tf32 / fp32) against a literal torch restatement --
  gather, Linear, activation, scatter_max with the first-maximum rule, autograd -- in both tensor-core modes and on both GEMM
  kernels (single CTA / CTA pair), on graphs with isolated pulses (deg 0), short events (deg < 8) and duplicate-heavy events
  (deg 9: torch_cluster's k + 1 quirk);
* LeakyReLU on random floats against the same restatement in fp64: outputs / gradients within the mode's stated tolerance."""

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _graph_and_inputs(seed, hid, c_out, integer, sizes=(1, 5, 9, 40, 1, 137, 300, 8, 2)):
    from graphnet_b200 import ops
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    # coarse grid => duplicate positions => some nodes keep k + 1 = 9 edges
    xyz = torch.randint(0, 3, (n, 3), generator=g).float()
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    graph = ops.knn_table(xyz.cuda(), [0, 1, 2], ptr.cuda(), 8)
    if integer:
        pq = torch.randint(-6, 7, (n, 2 * hid), generator=g).float()
        w2 = torch.randint(-2, 3, (c_out, hid), generator=g).float()
        b2 = torch.randint(-3, 4, (c_out,), generator=g).float()
        gy = torch.randint(-4, 5, (n, c_out), generator=g).float()
    else:
        pq = torch.randn(n, 2 * hid, generator=g)
        w2 = torch.randn(c_out, hid, generator=g) / hid ** 0.5
        b2 = torch.randn(c_out, generator=g) * 0.1
        gy = torch.randn(n, c_out, generator=g)
    return graph, pq, w2, b2, gy


def _act(v, code):
    return v if code == 0 else (torch.relu(v) if code == 1 else torch.nn.functional.leaky_relu(v, 0.01))


def _reference(graph, pq, w2, b2, gy, act1, act2, dtype, forced=None):
    """Literal restatement: per-slot messages, -inf on padding slots, maximum with the FIRST arg-max, gradient to that slot.
    forced = (slot, pos) [n, C]: the backward pass runs on these decisions (winning slot, sign of its pre-activation) instead
    of the restatement's own -- one decision within rounding of a tie / of the activation's kink otherwise moves whole
    gradient entries by O(1 / n), which says nothing about the kernels' arithmetic."""
    nbr, deg = graph.nbr.cpu().long(), graph.deg.cpu().long()
    n, width = nbr.shape
    hid = pq.shape[1] // 2
    pq = pq.to(dtype).requires_grad_(True)
    w2 = w2.to(dtype).requires_grad_(True)
    b2 = b2.to(dtype).requires_grad_(True)
    valid = torch.arange(width)[None, :] < deg[:, None]
    src = torch.where(valid, nbr, torch.zeros_like(nbr))
    h = _act(pq[:, None, :hid] + pq[src][:, :, hid:], act1)               # [n, W, H]
    pre = h @ w2.t() + b2
    m = _act(pre, act2)
    m_masked = torch.where(valid[:, :, None], m, torch.full_like(m, float("-inf")))
    top = m_masked.max(dim=1).values
    slots = torch.arange(width)[None, :, None].expand_as(m)
    first = torch.where(m_masked == top[:, None, :], slots, torch.full_like(slots, 99)).min(dim=1).values     # first maximum
    has = deg > 0
    pick = torch.gather(m, 1, first.clamp(max=width - 1)[:, None, :]).squeeze(1)
    y = torch.where(has[:, None], pick, torch.zeros_like(pick))
    pos = torch.gather(pre.detach(), 1, first.clamp(max=width - 1)[:, None, :]).squeeze(1) > 0
    if forced is None:
        (y * gy.to(dtype)).sum().backward()
    else:
        slot_k, pos_k = forced
        slope = {0: 1.0, 1: 0.0, 2: 0.01}[act2]
        pre_k = torch.gather(pre, 1, slot_k.clamp(min=0)[:, None, :]).squeeze(1)
        coef = gy.to(dtype) * torch.where(pos_k, torch.ones_like(pre_k), torch.full_like(pre_k, slope)) * has[:, None].to(dtype)
        (pre_k * coef.detach()).sum().backward()
    return y.detach(), first, pos, has, pq.grad, w2.grad, b2.grad, m.detach()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("acts", [(1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("hid,c_out", [(64, 256), (40, 72), (128, 128)])
def test_aggmax_route_bit_exact_on_integers(built_library, variant, precision, acts, hid, c_out):
    from graphnet_b200 import ops
    graph, pq, w2, b2, gy = _graph_and_inputs(7 + hid, hid, c_out, integer=True)
    assert int(graph.deg.max()) == 9 and int(graph.deg.min()) == 0, (int(graph.deg.max()), int(graph.deg.min()))   # k + 1 quirk, isolated pulse
    y_ref, first, pos, has, dpq_ref, dw_ref, db_ref, _ = _reference(graph, pq, w2, b2, gy, acts[0], acts[1], torch.float64)
    old = ops.PRECISION
    ops.set_precision(precision)
    ops._call("gnb_linear_set_variant", variant)
    try:
        pq_d, w_d, b_d = pq.cuda().requires_grad_(True), w2.cuda().requires_grad_(True), b2.cuda().requires_grad_(True)
        y = ops.edgeconv_hoisted_max(pq_d, w_d, b_d, graph, acts[0], acts[1])
        (y * gy.cuda()).sum().backward()
    finally:
        ops._call("gnb_linear_set_variant", 0)
        ops.set_precision(old)
    assert torch.equal(y.detach().cpu().double(), y_ref)
    assert torch.equal(pq_d.grad.cpu().double(), dpq_ref)
    assert torch.equal(w_d.grad.cpu().double(), dw_ref)
    assert torch.equal(b_d.grad.cpu().double(), db_ref)


def test_aggmax_arg_bytes(built_library):
    """The side output itself: slot of the first maximum, bit 0x40 = winning pre-activation > 0, -1 without neighbours."""
    from graphnet_b200 import ops
    hid, c_out = 64, 96
    graph, pq, w2, b2, gy = _graph_and_inputs(3, hid, c_out, integer=True)
    _, first, pos, has, *_ = _reference(graph, pq, w2, b2, gy, 1, 2, torch.float64)
    old = ops.PRECISION
    ops.set_precision("tf32")
    try:
        h = torch.empty(graph.n * 9, hid, device="cuda")
        pq_d, b_d = pq.cuda(), b2.cuda()
        ops._call("gnb_edge_hidden_fwd", ops._ptr(pq_d), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, graph.n,
                  ops.ACT_RELU | ops.FLAG_ROUND_TF32, ops._ptr(h), hid, ops._stream())
        w2p = ops._tc_pack_weight(w2.cuda(), (0,), (hid,))
        y = torch.empty(graph.n, c_out, device="cuda")
        arg = torch.full((graph.n, c_out), 77, dtype=torch.int8, device="cuda")
        ops._call("gnb_edge_linear_aggmax_fwd_tf32", ops._ptr(h), hid, hid, ops._ptr(w2p), w2p.shape[1], ops._ptr(b_d),
                  ops._ptr(graph.deg), graph.n, c_out, ops.ACT_LEAKY, 0, ops._ptr(y), c_out, ops._ptr(arg), c_out, ops._stream())
    finally:
        ops.set_precision(old)
    arg = arg.cpu().long()
    assert torch.equal(arg[~has], torch.full_like(arg[~has], -1))
    assert torch.equal(arg[has] & 0x3f, first[has])
    assert torch.equal((arg[has] & 0x40) != 0, pos[has])


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("tf32x3", 2e-5, 1e-3), ("tf32", 1e-3, 3e-3)])
def test_aggmax_route_leaky_relu_vs_fp64(built_library, precision, tol_out, tol_grad):
    """LeakyReLU twice (DynTrans' MLP) on random floats, graph of 4 000 pulses (a million arg-max decisions): outputs within
    the mode's stated tolerance; the kernel's decisions differ from the fp64 restatement's only between slots whose values
    agree within that tolerance; gradients, on the kernel's decisions, within the mode's stated gradient tolerance."""
    from graphnet_b200 import ops
    sizes = (700, 1, 1500, 64, 3, 1732)
    graph, pq, w2, b2, gy = _graph_and_inputs(11, 256, 256, integer=False, sizes=sizes)
    old = ops.PRECISION
    ops.set_precision(precision)
    try:
        pq_d, w_d, b_d = pq.cuda().requires_grad_(True), w2.cuda().requires_grad_(True), b2.cuda().requires_grad_(True)
        y = ops.edgeconv_hoisted_max(pq_d, w_d, b_d, graph, ops.ACT_LEAKY, ops.ACT_LEAKY)
        arg = y.grad_fn.saved_tensors[1].cpu().long()
        (y * gy.cuda()).sum().backward()
    finally:
        ops.set_precision(old)
    slot_k, pos_k = torch.where(arg >= 0, arg & 0x3f, torch.zeros_like(arg)), (arg >= 0) & ((arg & 0x40) != 0)
    y_ref, first, pos, has, dpq_ref, dw_ref, db_ref, m = _reference(graph, pq, w2, b2, gy, 2, 2, torch.float64, forced=(slot_k, pos_k))
    scale = float(y_ref.abs().max())
    m_first = torch.gather(m, 1, first.clamp(max=8)[:, None, :]).squeeze(1)
    m_kern = torch.gather(m, 1, slot_k[:, None, :]).squeeze(1)
    differ = has[:, None] & (slot_k != first)
    assert float(differ.float().mean()) < 1e-3
    assert float(((m_first - m_kern).abs() * differ).max()) < tol_out * scale                 # only near-ties flip
    errs = {"y": rel_err(y, y_ref), "dpq": rel_err(pq_d.grad, dpq_ref), "dw2": rel_err(w_d.grad, dw_ref),
            "db2": rel_err(b_d.grad, db_ref)}
    print(precision, errs, "decisions that differ:", int(differ.sum()))
    assert errs["y"] < tol_out, errs
    assert max(errs["dpq"], errs["dw2"], errs["db2"]) < tol_grad, errs


def test_encoder_layer_on_tokens_matches_the_padded_module_call(built_library):
    """DynTrans' per-event TransformerEncoder layer applied to packed tokens (four Linear layers on the tensor-core kernels,
    only q / k / v padded) against the literal `layer(dense, src_key_padding_mask=~mask)[mask]` of layers.py:190-195 in fp64."""
    import copy

    from graphnet_b200 import ops
    from graphnet_b200.models.components.layers import encoder_layer_on_tokens, to_dense_events
    torch.manual_seed(3)
    d, sizes = 64, (5, 1, 17, 40, 2)
    layer = torch.nn.TransformerEncoderLayer(d_model=d, nhead=4, dim_feedforward=2048, batch_first=True, norm_first=False).eval()
    ref = copy.deepcopy(layer).double()
    n = sum(sizes)
    x = torch.randn(n, d)
    w = torch.randn(n, d)
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    xr = x.double().requires_grad_(True)
    dense, mask = to_dense_events(xr, ptr)
    y_ref = ref(dense, src_key_padding_mask=~mask)[mask]
    (y_ref * w.double()).sum().backward()
    layer = layer.cuda()
    old = ops.PRECISION
    ops.set_precision("tf32x3")
    try:
        xd = x.cuda().requires_grad_(True)
        y = encoder_layer_on_tokens(layer, xd, ptr.cuda())
        (y * w.cuda()).sum().backward()
    finally:
        ops.set_precision(old)
    assert rel_err(y, y_ref) < 2e-5
    errs = {"x": rel_err(xd.grad, xr.grad)}
    for (k, p), (_, q) in zip(layer.named_parameters(), ref.named_parameters()):
        errs[k] = rel_err(p.grad, q.grad)
    print(errs)
    assert max(errs.values()) < 1e-3, errs
