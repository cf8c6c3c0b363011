"""tcgen05 (kind::tf32) Linear kernel and the tf32 precision mode of DynEdge.

Integer-valued operands are exactly representable in tf32 and their products/sums in fp32, so the tensor
core result must be BIT-EXACT there (this pins descriptors, swizzle, TMA boxes and the epilogue mapping).
Real-valued operands: rel 2e-3 per GEMM (tf32 has 10 mantissa bits: unit round-off 2^-11 per operand)."""

import numpy as np
import pytest
import torch

from helpers import namespace, oracle_on_kernel_decisions, rel_err
from oracle.dynedge_oracle import DynEdgeRef, batch_to_ptr, knn_graph_ref

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tf32_mode():
    from graphnet_b200 import ops
    old = ops.PRECISION
    ops.set_precision("tf32")
    yield ops
    ops.set_precision(old)


@pytest.fixture(params=[(1, 0), (2, 0), (2, 2), (3, 0)],
                ids=["single_cta", "cta_pair", "cta_pair_resident_weights", "cta_pair_dual_group_scatter"])
def linear_variant(request, built_library):
    """Run the tf32 Linear entry points on the single-CTA kernel and on the cta_group::2 CTA-pair kernel (weights
    streamed / resident in shared memory)."""
    from graphnet_b200 import _lib
    lib = _lib.load()
    variant, resident = request.param
    assert lib.gnb_linear_set_variant(variant) == 0
    assert lib.gnb_linear_set_pair_resident(resident) == 0
    yield request.param
    lib.gnb_linear_set_variant(0)
    lib.gnb_linear_set_pair_resident(0)


SHAPES = [  # rows, n_out, part widths
    (128, 128, [32]), (1, 1, [4]), (300, 336, [256]), (1000, 672, [32]), (515, 256, [336]),
    (777, 336, [32, 256, 256, 256, 256]), (129, 19, [20, 7]), (4099, 128, [1024]), (64, 700, [96, 40]),
]


@pytest.mark.parametrize("rows,n_out,widths", SHAPES)
def test_tc_linear_bit_exact_on_integers(built_library, tf32_mode, linear_variant, rows, n_out, widths):
    ops = tf32_mode
    g = torch.Generator().manual_seed(rows + n_out)
    parts = [torch.randint(-2, 3, (rows, w), generator=g).float() for w in widths]
    pw = [((w + 3) // 4) * 4 for w in widths]                      # caller convention: 4-aligned part offsets
    offsets = [int(v) for v in np.cumsum([0] + pw[:-1])]
    wmat = torch.zeros(n_out, sum(pw))
    for off, w in zip(offsets, widths):
        wmat[:, off:off + w] = torch.randint(-2, 3, (n_out, w), generator=g).float()
    bias = torch.randint(-3, 4, (n_out,), generator=g).float()
    ref = torch.zeros(rows, n_out, dtype=torch.float64)
    for p, off, w in zip(parts, offsets, widths):
        ref += p.double() @ wmat[:, off:off + w].double().t()
    ref = torch.relu(ref + bias.double())
    parts_padded = [torch.nn.functional.pad(p, (0, q - p.shape[1])) for p, q in zip(parts, pw)]
    out = ops.multi_linear_act([p.cuda() for p in parts_padded], wmat.cuda(), bias.cuda(), offsets, ops.ACT_RELU)
    assert torch.equal(out.cpu().double(), ref)


@pytest.mark.parametrize("rows,n_out,widths", SHAPES[2:6])
def test_tc_linear_forward_backward_real_valued(built_library, tf32_mode, rows, n_out, widths):
    ops = tf32_mode
    torch.manual_seed(rows)
    k = sum(widths)
    x = torch.randn(rows, k)
    w = torch.randn(n_out, k) / k ** 0.5
    b = torch.randn(n_out)
    gout = torch.randn(rows, n_out)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    offsets = [int(v) for v in np.cumsum([0] + widths[:-1])]
    xparts = [x[:, o:o + wd].contiguous().cuda().requires_grad_(True) for o, wd in zip(offsets, widths)]
    wg, bg = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.multi_linear_act(xparts, wg, bg, offsets, ops.ACT_RELU)
    (out * gout.cuda()).sum().backward()
    assert rel_err(out, ref) < 2e-3
    # gradients: the ReLU mask is taken from the kernel's own output (a pre-activation within rounding of 0
    # may legitimately land on either side in tf32), everything else in fp64
    dz = gout.double() * (out.detach().cpu() > 0)
    gx = torch.cat([p.grad for p in xparts], dim=1)
    assert rel_err(gx, dz @ w.double()) < 2e-3
    assert rel_err(wg.grad, dz.t() @ x.double()) < 2e-3
    assert rel_err(bg.grad, dz.sum(0)) < 2e-3


def test_dynedge_tf32_mode_vs_oracle(built_library, tf32_mode):
    """Default DynEdge in tf32 mode against the fp32 oracle fed the kernel's own graphs. Stated tolerance:
    rel 1e-3 on the model output (north star) and 3e-3 on parameter gradients; kNN graphs bit-exact on the
    kernel's own features."""
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(24, seed=5, n_max=400)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    kwargs = dict(global_pooling_schemes=["min", "max", "mean", "sum"])
    torch.manual_seed(0)
    ref = DynEdgeRef(7, **kwargs)
    model = DynEdge(7, **kwargs)
    model.load_state_dict(ref.state_dict())
    model = model.cuda()
    model._debug_record = True
    data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
    y = model(data)
    y.square().sum().backward()
    ptr = batch_to_ptr(batch)
    ei0 = knn_graph_ref(x[:, :3], 8, ptr=ptr)
    forced = [None]
    for li in range(1, 4):
        feats = model._debug["skips"][li].detach().cpu()
        ei_k = model._debug["graphs"][li].edge_index().cpu()
        assert torch.equal(ei_k, knn_graph_ref(feats[:, :3], 8, ptr=ptr)), f"latent graph {li}"
        forced.append(ei_k)
    y_ref, inter, _ = oracle_on_kernel_decisions(ref, namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses), forced, y, "tf32")
    y_ref.square().sum().backward()
    errs = {f"skip{li}": rel_err(model._debug["skips"][li], inter["skips"][li]) for li in range(5)}
    errs["out"] = rel_err(y, y_ref)
    gerr = {k: rel_err(p.grad, q.grad) for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters())}
    print("tf32 rel errors:", {k: f"{v:.2e}" for k, v in errs.items()}, "max grad", f"{max(gerr.values()):.2e}")
    print("tf32 grad rel errors:", {k: f"{v:.1e}" for k, v in gerr.items()})
    assert errs["out"] < 1e-3, errs
    assert max(gerr.values()) < 3e-3, gerr


WGRAD_SHAPES = [(64, 128, 32), (300, 256, 336), (1000, 336, 256), (4099, 672, 32), (515, 19, 20), (9000, 256, 256)]


@pytest.mark.parametrize("rows,n_out,k_in", WGRAD_SHAPES)
def test_tc_wgrad_bit_exact_on_integers(built_library, tf32_mode, rows, n_out, k_in):
    """dW += dz^T x with TMA-fed MN-major operands: exact on small integers (pins the MN-major descriptors)."""
    ops = tf32_mode
    g = torch.Generator().manual_seed(rows)
    ldz, ldx = ((n_out + 3) // 4) * 4, ((k_in + 3) // 4) * 4
    dz = torch.zeros(rows, ldz)
    x = torch.zeros(rows, ldx)
    dz[:, :n_out] = torch.randint(-2, 3, (rows, n_out), generator=g).float()
    x[:, :k_in] = torch.randint(-2, 3, (rows, k_in), generator=g).float()
    ref = dz[:, :n_out].double().t() @ x[:, :k_in].double()
    dzc, xc = dz.cuda()[:, :n_out], x.cuda()[:, :k_in]
    dw = torch.ones(n_out, k_in, device="cuda")                  # accumulate-onto semantics
    ops._gemm_bwd_weight_tc(dzc, xc, dw, k_in)
    assert torch.equal(dw.cpu().double(), ref + 1.0)


def _edgeconv_reference(pq, w2, b2, nbr, deg, aggr):
    """fp64 reference of y_i = AGG_s relu(W2 relu(P_i + Q_j) + b2) on the neighbour table."""
    n, width = nbr.shape
    h = pq.shape[1] // 2
    out = torch.zeros(n, w2.shape[0], dtype=torch.float64)
    for i in range(n):
        d = int(deg[i])
        if d == 0:
            continue
        j = nbr[i, :d].long()
        hid = torch.relu(pq[i, :h].double().unsqueeze(0) + pq[j, h:].double())
        msg = torch.relu(hid @ w2.double().t() + b2.double())
        out[i] = msg.sum(0) if aggr == "add" else msg.mean(0)
    return out


@pytest.mark.parametrize("hdim,c_out,k", [(336, 256, 8), (128, 256, 8), (32, 48, 4), (352, 128, 16), (64, 300, 8)])
@pytest.mark.parametrize("aggr", ["add", "mean"])
@pytest.mark.parametrize("variant", [1, 2], ids=["single_cta", "cta_pair"])
def test_fused_edgeconv_bit_exact_on_integers(built_library, tf32_mode, hdim, c_out, k, aggr, variant):
    """Fused gather + hidden ReLU + tcgen05 contraction + bias/ReLU + aggregation: exact on small integers
    (sum aggregation; mean compared to 1e-6) including degree-0, short and k+1-degree nodes."""
    ops = tf32_mode
    from helpers import tie_heavy_events
    from oracle.dynedge_oracle import batch_to_ptr
    sizes = [1, 2, 5, 9, 10, 64, 130, 12, 300]
    x, batch, _ = tie_heavy_events(sizes, 5, seed=hdim + k)
    x[-12:] = x[-12]
    ptr = batch_to_ptr(batch)
    graph = ops.knn_table(x.cuda(), [0, 1, 2], ptr.cuda(), k)
    n = x.shape[0]
    g = torch.Generator().manual_seed(c_out)
    pq = torch.randint(-2, 3, (n, 2 * hdim), generator=g).float()
    w2 = torch.randint(-1, 2, (c_out, hdim), generator=g).float()
    b2 = torch.randint(-3, 4, (c_out,), generator=g).float()
    ref = _edgeconv_reference(pq, w2, b2, graph.nbr.cpu(), graph.deg.cpu(), aggr)
    ops.set_edgeconv_variant(variant)      # the pair kernel only applies to 128 < c_out <= 256
    try:
        with torch.no_grad():
            y = ops.edgeconv_fused_forward(pq.cuda(), w2.cuda(), b2.cuda(), graph, aggr)
        torch.cuda.synchronize()
    finally:
        ops.set_edgeconv_variant(0)
    if aggr == "add":
        assert torch.equal(y.cpu().double(), ref)
    else:
        assert rel_err(y, ref) < 1e-3          # output is rounded to tf32 (unit round-off 2^-11)


def test_fused_edgeconv_matches_unfused_route_and_executor(built_library, tf32_mode):
    ops = tf32_mode
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(20, seed=12, n_max=500)
    x, batch, n_pulses = (torch.from_numpy(raw[k]).cuda() for k in ("x", "batch", "n_pulses"))
    torch.manual_seed(2)
    model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
    outs = {}
    old, old_route = ops.FUSED_EDGECONV, ops.INFERENCE_ROUTE
    try:
        # per-kernel route, hidden layer + aggregating GEMM epilogue ("split"), single fused EdgeConv kernel ("fused")
        for name, fused, route in (("unfused", False, "split"), ("split", True, "split"), ("fused", True, "fused")):
            ops.FUSED_EDGECONV, ops.INFERENCE_ROUTE = fused, route
            with torch.no_grad():
                outs[name] = model(KNNEdges(8)(Data(x=x, batch=batch, n_pulses=n_pulses))).clone()
    finally:
        ops.FUSED_EDGECONV, ops.INFERENCE_ROUTE = old, old_route
    # same rounding points except the per-edge message (not rounded to tf32 before the k-sum in the fused kernels);
    # the latent kNN graphs of the runs may differ where rounding flips a near-tie, hence the looser bound
    assert rel_err(outs["split"], outs["unfused"]) < 5e-3
    assert rel_err(outs["fused"], outs["unfused"]) < 5e-3


@pytest.mark.parametrize("route", ["split", "fused"])
def test_dynedge_tf32_fused_inference_vs_oracle(built_library, tf32_mode, route, monkeypatch):
    """Inference paths (hidden layer + aggregating tcgen05 GEMM; single fused tcgen05 EdgeConv kernel) against the fp32
    oracle fed the kernel's own graphs: rel 1e-3."""
    from graphnet_b200 import Data
    monkeypatch.setattr(tf32_mode, "INFERENCE_ROUTE", route)
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(24, seed=5, n_max=400)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    kwargs = dict(global_pooling_schemes=["min", "max", "mean", "sum"])
    torch.manual_seed(0)
    ref = DynEdgeRef(7, **kwargs)
    model = DynEdge(7, **kwargs)
    model.load_state_dict(ref.state_dict())
    model = model.cuda()
    model._debug_record = True
    with torch.no_grad():
        y = model(KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda())))
    ptr = batch_to_ptr(batch)
    forced = [None]
    for li in range(1, 4):
        feats = model._debug["skips"][li].detach().cpu()
        ei_k = model._debug["graphs"][li].edge_index().cpu()
        assert torch.equal(ei_k, knn_graph_ref(feats[:, :3], 8, ptr=ptr)), f"latent graph {li}"
        forced.append(ei_k)
    ei0 = knn_graph_ref(x[:, :3], 8, ptr=ptr)
    with torch.no_grad():
        y_ref = ref(namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses), forced_graphs=forced)
    err = rel_err(y, y_ref)
    print("tf32 fused inference rel error:", f"{err:.2e}")
    assert err < 1e-3


@pytest.mark.parametrize("hdim,c_out", [(336, 256), (128, 256), (64, 40), (512, 96)])
def test_edge_hidden_dgrad_scatter_bit_exact_on_integers(built_library, tf32_mode, linear_variant, hdim, c_out):
    """Hidden layer forward with activation bit mask, then dgrad GEMM + ReLU mask + scatter epilogue: dP = sum over
    slots, dQ scattered to the neighbours. Integer operands make every product and (order-independent) sum exact, so
    the result must equal the fp64 reference."""
    ops = tf32_mode
    from helpers import tie_heavy_events
    sizes = [1, 2, 5, 9, 10, 64, 130, 12, 300]
    x, batch, _ = tie_heavy_events(sizes, 5, seed=hdim)
    x[-12:] = x[-12]                                   # duplicates: degree k + 1
    ptr = batch_to_ptr(batch)
    graph = ops.knn_table(x.cuda(), [0, 1, 2], ptr.cuda(), 8)
    n = x.shape[0]
    nbr, deg = graph.nbr.cpu().long(), graph.deg.cpu()
    assert nbr.shape[1] == 9
    g = torch.Generator().manual_seed(c_out)
    pq = torch.randint(-2, 3, (n, 2 * hdim), generator=g).float()
    dz = torch.randint(-2, 3, (n * 9, c_out), generator=g).float()
    w2 = torch.randint(-1, 2, (c_out, hdim), generator=g).float()
    # reference: h[(i,s)] = relu(P_i + Q_j); dh = (dz W2) * (h > 0); scatter
    h_ref = torch.zeros(n * 9, hdim, dtype=torch.float64)
    for i in range(n):
        for s_ in range(int(deg[i])):
            h_ref[i * 9 + s_] = torch.relu(pq[i, :hdim] + pq[nbr[i, s_], hdim:]).double()
    dh = (dz.double() @ w2.double()) * (h_ref > 0).double()
    dpq_ref = torch.zeros(n, 2 * hdim, dtype=torch.float64)
    for i in range(n):
        for s_ in range(int(deg[i])):
            r = i * 9 + s_
            dpq_ref[i, :hdim] += dh[r]
            dpq_ref[nbr[i, s_], hdim:] += dh[r]
    # device
    mld = 4 * ((hdim + 127) // 128)
    ntile = (n + 13) // 14
    hmask = torch.full((ntile * 126, mld), -1, dtype=torch.int32, device="cuda")     # garbage in the padding rows
    hdev = torch.empty(n * 9, hdim, device="cuda")
    pqc = pq.cuda()
    ops._call("gnb_edge_hidden_fwd_mask", ops._ptr(pqc), 2 * hdim, hdim, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n,
              1 | 0x100, ops._ptr(hdev), hdim, ops._ptr(hmask), mld, ops._stream())
    assert torch.equal(hdev.cpu().double(), h_ref)
    bits = hmask[: n * 9].cpu().numpy().view(np.uint32)
    expect = np.zeros((n * 9, mld), dtype=np.uint32)
    hpos = (h_ref > 0).numpy()
    for c in range(hdim):      # layout: channel c = bit (c % 128) / 4 of word 4 (c / 128) + c % 4
        expect[:, 4 * (c // 128) + c % 4] |= hpos[:, c].astype(np.uint32) << np.uint32((c % 128) // 4)
    assert np.array_equal(bits, expect)
    kpad = (c_out + 31) // 32 * 32
    wt = torch.zeros(hdim, kpad)
    wt[:, :c_out] = w2.t()
    dzc, wtc = dz.cuda(), wt.cuda()
    dpq = torch.zeros(n, 2 * hdim, device="cuda")
    dpq[:, :hdim] = 7.0                                # the P half is overwritten, not accumulated
    ops._call("gnb_edge_hidden_dgrad_scatter_tf32", ops._ptr(dzc), c_out, c_out, ops._ptr(wtc), kpad, ops._ptr(hmask), mld,
              hdim, ops._ptr(graph.nbr), n, ops._ptr(dpq), 2 * hdim, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dpq.cpu().double(), dpq_ref)
    # split form used by the step executor: Q half reduced into its own tensor (odd pitch), P half written to a second
    # tensor (tf32 rounding is exact on these integers) and the P half's column sums added onto a bias-gradient vector
    dq = torch.zeros(n, hdim + 8, device="cuda")
    dp = torch.full((n, hdim + 4), 7.0, device="cuda")
    dbias = torch.full((hdim,), 3.0, device="cuda")
    ops._call("gnb_edge_hidden_dgrad_scatter_split_tf32", ops._ptr(dzc), c_out, c_out, ops._ptr(wtc), kpad, ops._ptr(hmask), mld,
              hdim, ops._ptr(graph.nbr), n, ops._ptr(dq), hdim + 8, ops._ptr(dp), hdim + 4, ops._ptr(dbias), 0x100, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dq[:, :hdim].cpu().double(), dpq_ref[:, hdim:])
    assert torch.equal(dq[:, hdim:].cpu(), torch.zeros(n, 8))
    assert torch.equal(dp[:, :hdim].cpu().double(), dpq_ref[:, :hdim])
    assert torch.equal(dp[:, hdim:].cpu(), torch.full((n, 4), 7.0))
    assert torch.equal(dbias.cpu().double(), dpq_ref[:, :hdim].sum(0) + 3.0)


def test_act_bwd_colsum_zero_source_flag(built_library, tf32_mode):
    """gnb_act_bwd_colsum with flag 0x400 on a strided half of an accumulation buffer: the rounded copy is written, the
    source half is left zeroed and the other half untouched (how the executor recycles the dQ accumulation buffer)."""
    ops = tf32_mode
    g = torch.Generator().manual_seed(5)
    n, hdim = 1000, 336
    acc = torch.randn(n, 2 * hdim, generator=g).cuda()
    keep = acc.clone()
    out = torch.full((n, 2 * hdim), -1.0, device="cuda")
    ops._call("gnb_act_bwd_colsum", acc.data_ptr() + 4 * hdim, 2 * hdim, None, 0, n, hdim, out.data_ptr() + 4 * hdim, 2 * hdim,
              None, 0x100 | 0x400, None, 1, 0, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(acc[:, :hdim], keep[:, :hdim]) and torch.equal(acc[:, hdim:], torch.zeros_like(acc[:, hdim:]))
    assert torch.equal(out[:, :hdim], torch.full_like(out[:, :hdim], -1.0))
    # cvt.rna.tf32.f32: nearest, ties away from zero, 10 mantissa bits = add half an ulp to the magnitude bits, truncate
    bits = keep[:, hdim:].cpu().numpy().view(np.uint32)
    expect = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    assert np.array_equal(out[:, hdim:].cpu().numpy(), expect)


@pytest.mark.parametrize("k,n_out", [(336, 256), (128, 256), (40, 100), (352, 336)])
def test_edge_linear_agg_bit_exact_on_integers(built_library, tf32_mode, linear_variant, k, n_out):
    """Second EdgeConv Linear + ReLU + k-sum + mask bits in the GEMM epilogue, and the mask-bit backward kernel."""
    ops = tf32_mode
    from helpers import tie_heavy_events
    sizes = [1, 2, 5, 9, 10, 64, 130, 12, 300, 3]
    x, batch, _ = tie_heavy_events(sizes, 5, seed=k)
    x[-15:-3] = x[-15]
    ptr = batch_to_ptr(batch)
    graph = ops.knn_table(x.cuda(), [0, 1, 2], ptr.cuda(), 8)
    n = x.shape[0]
    deg = graph.deg.cpu()
    g = torch.Generator().manual_seed(n_out)
    h = torch.randint(-1, 3, (n * 9, k), generator=g).float()
    w = torch.randint(-1, 2, (n_out, k), generator=g).float()
    b = torch.randint(-3, 4, (n_out,), generator=g).float()
    pre = h.double() @ w.double().t() + b.double()
    valid = (torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)).reshape(-1)            # [n*9]
    on = (pre > 0) & valid.unsqueeze(1)
    y_ref = (pre * on).reshape(n, 9, n_out).sum(1)
    kpad = (k + 31) // 32 * 32
    wp = torch.zeros(n_out, kpad)
    wp[:, :k] = w
    hc, wpc, bc = h.cuda(), wp.cuda(), b.cuda()
    y = torch.empty(n, n_out, device="cuda")
    ntile = (n + 13) // 14
    mask = torch.zeros(ntile * n_out * 4, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_linear_agg_fwd_tf32", ops._ptr(hc), k, k, ops._ptr(wpc), kpad, ops._ptr(bc), ops._ptr(graph.deg), n,
              n_out, 1, ops._ptr(y), n_out, ops._ptr(mask), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y.cpu().double(), y_ref)
    # backward from the mask bits: dz[(i,s)] = g[i] * bit
    gy = torch.randint(-2, 3, (n, n_out), generator=g).float()
    dz = torch.full((n * 9, n_out), 7.0, device="cuda")
    db = torch.zeros(n_out, device="cuda")
    gyc = gy.cuda()
    ops._call("gnb_edge_mask_bwd_colsum", ops._ptr(gyc), n_out, ops._ptr(mask), n, n_out, ops._ptr(graph.deg), ops._ptr(dz),
              n_out, ops._ptr(db), 0x100, ops._stream())
    torch.cuda.synchronize()
    dz_ref = gy.double().repeat_interleave(9, dim=0) * on
    assert torch.equal(dz.cpu().double(), dz_ref)
    assert torch.equal(db.cpu().double(), dz_ref.sum(0))


@pytest.mark.parametrize("rows,n_out,k", [(70000, 256, 336), (515, 336, 256), (129, 19, 40)])
def test_tc_linear_accumulate_flag(built_library, tf32_mode, linear_variant, rows, n_out, k):
    """act | 0x200: y += x W^T (the data-gradient accumulation onto the skip path), bit-exact on integers."""
    import ctypes
    ops = tf32_mode
    g = torch.Generator().manual_seed(rows)
    x = torch.randint(-2, 3, (rows, k), generator=g).float()
    w = torch.randint(-2, 3, (n_out, k), generator=g).float()
    y0 = torch.randint(-5, 6, (rows, n_out), generator=g).float()
    ref = y0.double() + x.double() @ w.double().t()
    kp = (k + 3) // 4 * 4
    xc = torch.nn.functional.pad(x, (0, kp - k)).cuda()
    packed = ops._tc_pack_weight(torch.nn.functional.pad(w, (0, kp - k)).cuda(), (0,), (kp,))
    y = y0.cuda()
    xs = (ctypes.c_void_p * 1)(xc.data_ptr())
    lds = (ctypes.c_int64 * 1)(kp)
    ks = (ctypes.c_int32 * 1)(kp)
    ops._call("gnb_linear_fwd_tf32", xs, lds, ks, 1, ops._ptr(packed), packed.shape[1], ops._ptr(None), ops._ptr(y), n_out,
              rows, n_out, 0x200, 0, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y.cpu().double(), ref)


@pytest.mark.parametrize("variant", [2, 3], ids=["cta_pair", "dual_group"])
@pytest.mark.parametrize("hdim,c_out,n", [(336, 256, 7013), (512, 96, 9001), (128, 256, 6500)])
def test_edge_hidden_dgrad_scatter_many_tiles_per_cluster(built_library, tf32_mode, variant, hdim, c_out, n):
    """The persistent pipelines with SEVERAL row tiles per cluster (barrier phase wrap-around, slot / metadata / TMEM buffer
    reuse) -- the small bit-exact case above gives every cluster at most one tile. Random integer operands, random
    activation bits, random in-event neighbours (some nodes with fewer than 8 and some with 9 neighbours); the fp64
    reference is vectorised. Exact equality (products and order-independent sums are exact on these integers)."""
    ops = tf32_mode
    from graphnet_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(hdim + n)
    rows = n * 9
    nbr = torch.randint(0, n, (n, 9), generator=g)
    base = torch.arange(n).unsqueeze(1)
    nbr = (base + torch.randint(-40, 41, (n, 9), generator=g)).clamp_(0, n - 1)
    deg = torch.full((n,), 8, dtype=torch.int64)
    deg[torch.randint(0, n, (n // 20,), generator=g)] = 9                     # duplicate quirk: slot 8 holds an edge
    few = torch.randint(0, n, (n // 15,), generator=g)
    deg[few] = torch.randint(0, 8, (few.numel(),), generator=g)
    valid = torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)                   # [n, 9]
    nbr = torch.where(valid, nbr, torch.full_like(nbr, -1))
    dz = torch.randint(-2, 3, (rows, c_out), generator=g).float()
    w2 = torch.randint(-1, 2, (c_out, hdim), generator=g).float()
    bits = (torch.rand(rows, hdim, generator=g) < 0.5) & valid.reshape(-1, 1)  # padding slots carry no activation bits
    mld = 4 * ((hdim + 127) // 128)
    ntile = (n + 13) // 14
    words = np.zeros((ntile * 126, mld), dtype=np.uint32)
    words[rows:] = 0xFFFFFFFF                                                  # garbage behind the last node
    bn = bits.numpy()
    for c in range(hdim):      # layout of gnb_edge_hidden_fwd_mask
        words[:rows, 4 * (c // 128) + c % 4] |= bn[:, c].astype(np.uint32) << np.uint32((c % 128) // 4)
    dh = (dz.double() @ w2.double()) * bits.double()
    dp_ref = dh.reshape(n, 9, hdim).sum(1)
    dq_ref = torch.zeros(n, hdim, dtype=torch.float64)
    flat_nbr = nbr.reshape(-1)
    sel = flat_nbr >= 0
    dq_ref.index_add_(0, flat_nbr[sel], dh[sel])
    kpad = (c_out + 31) // 32 * 32
    wt = torch.zeros(hdim, kpad)
    wt[:, :c_out] = w2.t()
    hmask = torch.from_numpy(words.view(np.int32)).cuda()
    dzc, wtc, nbrc = dz.cuda(), wt.cuda(), nbr.int().cuda()
    dq = torch.zeros(n, hdim, device="cuda")
    dp = torch.full((n, hdim), 7.0, device="cuda")
    dbias = torch.zeros(hdim, device="cuda")
    assert lib.gnb_linear_set_variant(variant) == 0
    try:
        ops._call("gnb_edge_hidden_dgrad_scatter_split_tf32", ops._ptr(dzc), c_out, c_out, ops._ptr(wtc), kpad, ops._ptr(hmask), mld,
                  hdim, ops._ptr(nbrc), n, ops._ptr(dq), hdim, ops._ptr(dp), hdim, ops._ptr(dbias), 0x100, ops._stream())
        torch.cuda.synchronize()
    finally:
        lib.gnb_linear_set_variant(0)
    assert torch.equal(dp.cpu().double(), dp_ref)
    assert torch.equal(dq.cpu().double(), dq_ref)
    assert torch.equal(dbias.cpu().double(), dp_ref.sum(0))
