"""The 8-slot edge layout of the fp16-plane per-edge kernels (include/graphnet_b200.h: gnb_edge_slot_flag and the four _w entry
points). The neighbour table is k + 1 = 9 wide only because a node with more than k exact duplicates keeps k + 1 edges
(torch_cluster's knn_graph behind edges.py:72-80 / layers.py:63-67); on a graph without such a node the device-side flag
switches the layer's four kernels to 16-node x 8-slot tiles. Every kernel is pinned bit-exactly on integer-valued operands
against fp64 / against its own 9-slot run, then the executor's training step is compared between the two layouts."""

import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle.dynedge_oracle import batch_to_ptr
from test_gpu_bf16 import _pq_lane_interleaved, _scale_word, f16_planes, two_plane_values

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ops(built_library):
    from graphnet_b200 import ops as _ops
    old = _ops.PRECISION
    yield _ops
    _ops.set_precision(old)
    _ops.SLOTS9 = False


def _graph8(ops, sizes, seed):
    """Random positions (no duplicates): every node has min(8, n_event - 1) neighbours, none has 9."""
    g = torch.Generator().manual_seed(seed)
    n = int(sum(sizes))
    x = torch.randn(n, 3, generator=g)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    graph = ops.knn_table(x.cuda(), [0, 1, 2], batch_to_ptr(batch).cuda(), 8)
    assert int(graph.deg.max()) <= 8
    return graph, n


def _flag(ops, graph, n):
    f = torch.full((1,), 77, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_slot_flag", ops._ptr(graph.deg), n, 8, ops._ptr(f), ops._stream())
    return f


def _i32(t):
    return torch.where(t >= 2 ** 31, t - 2 ** 32, t).int()


def test_slot_flag(ops):
    graph, n = _graph8(ops, [5, 300, 9, 1200], seed=1)
    assert int(_flag(ops, graph, n)) == 0
    deg = graph.deg.clone()
    deg[n - 1] = 9
    f = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_slot_flag", ops._ptr(deg), n, 8, ops._ptr(f), ops._stream())
    assert int(f) == 1
    ops._call("gnb_edge_slot_flag", ops._ptr(graph.deg), 0, 8, ops._ptr(f), ops._stream())          # empty graph: 8 slots
    assert int(f) == 0
    # the many-CTA form ORs into a word the caller zeroed
    ops._call("gnb_edge_slot_flag_or", ops._ptr(graph.deg), n, 8, ops._ptr(f), ops._stream())
    assert int(f) == 0
    ops._call("gnb_edge_slot_flag_or", ops._ptr(deg), n, 8, ops._ptr(f), ops._stream())
    assert int(f) == 1
    ops._call("gnb_edge_slot_flag_or", ops._ptr(graph.deg), n, 8, ops._ptr(f), ops._stream())        # never cleared by the kernel
    assert int(f) == 1


def _mask_case8(ops, sizes, n_out, seed, gmax):
    """Random ReLU bits in the 8-slot layout of the aggregating epilogue, a gradient g and the dz they imply; runs
    gnb_edge_dz_prep_w with a zero flag and checks its three outputs."""
    graph, n = _graph8(ops, sizes, seed)
    flag = _flag(ops, graph, n)
    gen = torch.Generator().manual_seed(seed)
    deg = graph.deg.cpu()
    valid = torch.arange(8).unsqueeze(0) < deg.unsqueeze(1)                              # [n, 8]
    bits = (torch.rand(n, 8, n_out, generator=gen) < 0.55) & valid.unsqueeze(2)          # [n, 8, C]
    ntile = (n + 15) // 16
    full = torch.zeros(ntile * 16, 8, n_out, dtype=torch.bool)
    full[:n] = bits
    cols = full.reshape(ntile, 128, n_out)                                                # bit index 8 f + s
    words = torch.zeros(ntile, n_out, 4, dtype=torch.int64)
    for w in range(4):
        words[:, :, w] = (cols[:, 32 * w: 32 * w + 32].long() << torch.arange(32).view(1, -1, 1)).sum(1)
    words = _i32(words).reshape(-1).cuda()
    rb = full.reshape(ntile * 128, n_out // 32, 32).long()
    rowm = _i32((rb << torch.arange(32).view(1, 1, 32)).sum(2))
    g = torch.randint(-4, 5, (n, n_out), generator=gen).float() * (gmax / 4)
    dz = (g.unsqueeze(1) * bits).reshape(n * 8, n_out)
    wz, sz = _scale_word(gmax)
    g16 = torch.full((n, n_out), 7.0, dtype=torch.float16, device="cuda")
    rows_alloc = max(ntile * 128, (n + 13) // 14 * 126)
    rowmask = torch.full((rows_alloc, n_out // 32), -1, dtype=torch.int32, device="cuda")
    db = torch.zeros(n_out, device="cuda")
    ops._call("gnb_edge_dz_prep_w", ops._ptr(g.cuda()), n_out, ops._ptr(words), n, n_out, ops._ptr(wz), ops._ptr(g16),
              ops._ptr(rowmask), ops._ptr(db), ops._ptr(flag), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(g16.cpu(), (g * sz).half())
    assert torch.equal(rowmask.cpu()[: ntile * 128], rowm)
    assert torch.equal(db.cpu().double(), dz.double().sum(0))
    return graph, n, flag, wz, g16, dz, rowmask


@pytest.mark.parametrize("n_out,k_in,sizes", [(256, 336, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700]), (256, 128, [400, 900, 14, 15, 16, 17]),
                                              (96, 344, [77, 5, 230]), (256, 336, [3000, 2500])])
def test_wgrad_f16_masked_8slot_bit_exact(ops, n_out, k_in, sizes):
    graph, n, flag, wz, g16, dz, rowm = _mask_case8(ops, sizes, n_out, seed=n_out + k_in, gmax=2.0 ** -9)
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(-2, 3, (n * 8, k_in), generator=gen).float()
    wx, sx = _scale_word(2.0)
    ref = dz.double().t() @ x.double()
    # x lives in a buffer sized for the 9-slot layout whose tail holds NaN patterns: rows beyond 8 n must never be read
    x16 = torch.full((n * 9, k_in), float("nan"), dtype=torch.float16, device="cuda")
    x16[: n * 8] = (x * sx).half().cuda()
    dw = torch.zeros(n_out, k_in, device="cuda")
    ops._call("gnb_linear_bwd_weight_f16_masked_w", ops._ptr(g16), ops._ptr(rowm), ops._ptr(x16), k_in, ops._ptr(dw), k_in, n,
              n_out, k_in, ops._ptr(wz), ops._ptr(wx), ops._ptr(flag), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dw.cpu().double(), ref)


@pytest.mark.parametrize("rowmajor", [False, True])
@pytest.mark.parametrize("hdim,c_out,sizes", [(336, 256, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 500]), (128, 256, [400, 900, 14, 15, 16, 17]),
                                              (40, 128, [77, 5, 230]), (336, 256, [3000, 2500]), (512, 64, [100, 37])])
def test_dgrad_scatter_f16_masked_8slot_bit_exact(ops, hdim, c_out, sizes, rowmajor):
    graph, n, flag, word, g16, dz, rowm = _mask_case8(ops, sizes, c_out, seed=hdim + c_out, gmax=2.0 ** -20)
    gen = torch.Generator().manual_seed(hdim)
    w2 = torch.randint(-8, 9, (c_out, hdim), generator=gen).float() / 8
    nbr, deg = graph.nbr.cpu().long()[:, :8], graph.deg.cpu()
    valid = ((torch.arange(8).unsqueeze(0) < deg.unsqueeze(1)) & (nbr >= 0)).reshape(-1)
    hbits = torch.rand(n * 8, hdim, generator=gen) < 0.6
    mld = 4 * ((hdim + 127) // 128)
    rows_m = max((n + 15) // 16 * 128, (n + 13) // 14 * 126)
    c = torch.arange(hdim)
    word_i, bit = (c // 32, c % 32) if rowmajor else (4 * (c // 128) + c % 4, (c % 128) // 4)
    hm = torch.zeros(rows_m, mld, dtype=torch.int64)
    hm[: n * 8].index_put_((torch.arange(n * 8).unsqueeze(1).expand(-1, hdim), word_i.unsqueeze(0).expand(n * 8, -1)),
                           hbits.long() << bit.unsqueeze(0), accumulate=True)
    hm = _i32(hm).cuda()
    da = (dz.double() @ w2.double()) * hbits
    dp_ref = da.reshape(n, 8, hdim).sum(1)
    dq_ref = torch.zeros(n, hdim, dtype=torch.float64)
    dq_ref.index_add_(0, nbr.reshape(-1).clamp(min=0)[valid], da[valid])
    cw = (c_out + 63) // 64 * 64
    wt16, _ = f16_planes(ops, w2.cuda(), 1, dst_cols=cw, transpose=True)
    dq = torch.zeros(n, hdim, device="cuda")
    dp = torch.full((n, hdim), 9.0, device="cuda")
    dbias = torch.zeros(hdim, device="cuda")
    ops._call("gnb_edge_hidden_dgrad_scatter_f16_masked_w", ops._ptr(g16), ops._ptr(rowm), c_out, ops._ptr(wt16), cw, ops._ptr(hm),
              mld, hdim, ops._ptr(graph.nbr), n, ops._ptr(dq), hdim, ops._ptr(dp), hdim, ops._ptr(dbias), 0x800 if rowmajor else 0,
              ops._ptr(word), ops._ptr(flag), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dp.cpu().double(), dp_ref)
    assert torch.equal(dq.cpu().double(), dq_ref)
    assert torch.equal(dbias.cpu().double(), dp_ref.sum(0))


def _decode_words(words, ntile, npt, w, n_out, n):
    """mask words [(tile, channel), 4] -> bits [n, w, C]"""
    wd = words.cpu().long().view(ntile, n_out, 4) & 0xFFFFFFFF
    bits = ((wd.unsqueeze(3) >> torch.arange(32).view(1, 1, 1, 32)) & 1).reshape(ntile, n_out, 128)[:, :, : npt * w]
    return bits.permute(0, 2, 1).reshape(ntile * npt, w, n_out)[:n].bool()


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("hid,n_out,sizes", [(336, 256, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700]), (128, 256, [400, 900, 14, 15, 16, 17]),
                                             (40, 104, [77, 5, 230]), (336, 256, [3000, 2500]), (128, 256, [16 * 5])])
def test_edgeconv_fused_fwd_f16_8slot_matches_9slot(ops, np_, hid, n_out, sizes, layout):
    """The same launch with the flag word 0 (8-slot tiles) and without a flag (9-slot tiles): y bit for bit (a node's eight
    messages are summed in the same order in both), plane 0 of h / the ReLU bytes / the mask words agree slot by slot."""
    graph, n = _graph8(ops, sizes, seed=hid + n_out)
    flag = _flag(ops, graph, n)
    assert int(flag) == 0
    gen = torch.Generator().manual_seed(hid)
    pq = torch.randint(-6, 7, (n, 2 * hid), generator=gen).float() * 0.25
    w = (torch.randint(-2, 3, (n_out, hid), generator=gen).float() if np_ == 1 else two_plane_values((n_out, hid), gen, scale_bits=6, mag_bits=13))
    b = torch.randint(-3, 4, (n_out,), generator=gen).float()
    pqc, bc = (_pq_lane_interleaved(pq, hid) if layout == 1 else pq).cuda(), b.cuda()
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops._call("gnb_absmax_bits", ops._ptr(pq.cuda()), 2 * hid, n, 2 * hid, 1, ops._ptr(word), ops._stream())
    kw = (hid + 63) // 64 * 64
    w0, w1 = f16_planes(ops, w.cuda(), np_, dst_cols=kw)
    mld = 4 * ((hid + 127) // 128)
    nt9, nt8 = (n + 13) // 14, (n + 15) // 16
    rows_alloc = max(nt9 * 126, nt8 * 128)
    out = {}
    for name, fl in (("w9", None), ("w8", flag)):
        y = torch.empty(n, n_out, device="cuda")
        m = torch.zeros(nt9 * n_out * 4, dtype=torch.int32, device="cuda")
        h0 = torch.full((n * 9, hid), 5.0, dtype=torch.float16, device="cuda")
        hb_all = torch.zeros(rows_alloc * mld * 4 + 4096, dtype=torch.uint8, device="cuda")
        hb_all[rows_alloc * mld * 4:] = 77
        ops._call("gnb_edgeconv_fused_fwd_f16_w", ops._ptr(pqc), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), n, ops._ptr(w0),
                  ops._ptr(w1), kw, ops._ptr(bc), n_out, 0, ops._ptr(y), n_out, ops._ptr(m), ops._ptr(h0), hid, ops._ptr(hb_all), mld * 4,
                  ops._ptr(word), layout, ops._ptr(fl), ops._stream())
        torch.cuda.synchronize()
        assert bool((hb_all[rows_alloc * mld * 4:] == 77).all())
        out[name] = (y, m, h0, hb_all[: rows_alloc * mld * 4].view(rows_alloc, mld * 4))
    y9, m9, h9, b9 = out["w9"]
    y8, m8, h8, b8 = out["w8"]
    assert torch.equal(y8, y9)
    assert torch.equal(h8[: n * 8].view(n, 8, hid), h9.view(n, 9, hid)[:, :8])
    assert bool((h8[n * 8:] == 5.0).all())                               # nothing written beyond the 8 n rows
    assert torch.equal(b8[: n * 8].view(n, 8, -1)[:, :, : hid // 8], b9[: n * 9].view(n, 9, -1)[:, :8, : hid // 8])
    bits9 = _decode_words(m9, nt9, 14, 9, n_out, n)
    bits8 = _decode_words(m8[: nt8 * n_out * 4], nt8, 16, 8, n_out, n)
    assert torch.equal(bits8, bits9[:, :8]) and not bits9[:, 8].any()
    # inference form: no side outputs
    y2 = torch.empty(n, n_out, device="cuda")
    ops._call("gnb_edgeconv_fused_fwd_f16_w", ops._ptr(pqc), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), n, ops._ptr(w0),
              ops._ptr(w1), kw, ops._ptr(bc), n_out, 0, ops._ptr(y2), n_out, ops._ptr(None), ops._ptr(None), hid, ops._ptr(None), mld * 4,
              ops._ptr(word), layout, ops._ptr(flag), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y2, y9)


@pytest.mark.parametrize("dup", [False, True])
def test_mixed16_training_step_8slot_layout_matches_9slot_layout(ops, dup):
    """Executor (mixed16): the default run (8-slot layout wherever the device flag allows it) against flags bit 4 (ops.SLOTS9:
    9 slots everywhere). Same products; a node's messages are summed in an order that depends on where its slot columns sit in
    the tile (the 9-slot tile sums a node that straddles two 32-column TMEM loads in two pieces), so the first layer agrees to
    fp32 summation order and later layers, whose kNN graphs are rebuilt from those features, at the mode's tolerance (a near-tie
    may pick another neighbour). dup = True plants duplicate positions: layer 1 then runs on 9 slots in both runs."""
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    ops.set_precision("mixed16")
    raw = make_batch(24, seed=11, n_max=500)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    if dup:
        x[40:52] = x[40]
    torch.manual_seed(3)
    model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
    data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
    model._debug_record = True
    outs, skips, grads = [], [], []
    for slots9 in (False, True):
        ops.SLOTS9 = slots9
        try:
            model.zero_grad(set_to_none=True)
            y = model(data)
            skips.append([s.detach().clone() for s in model._debug["skips"]])
            y.square().sum().backward()
            outs.append(y.detach().clone())
            grads.append([p.grad.clone() for p in model.parameters()])
        finally:
            ops.SLOTS9 = False
    assert rel_err(skips[0][1], skips[1][1]) < 2e-6
    if dup:
        assert torch.equal(skips[0][1], skips[1][1])                     # layer 1 ran the same 9-slot kernels twice
    assert rel_err(outs[0], outs[1]) < 1e-3
    for a, b in zip(*grads):
        assert rel_err(a, b) < 2e-2


@pytest.mark.parametrize("slots8", [False, True])
def test_dz_prep_zero_job(ops, slots8):
    """gnb_edge_dz_prep_wz: the zero job riding on the launch clears exactly the [rows, cols] block of pitch lda and the scratch
    vector, and the kernel's own outputs are those of the plain call."""
    graph, n, flag, wz, g16_ref, dz, rowmask_ref = _mask_case8(ops, [300, 5, 77, 1000], 256, seed=3, gmax=2.0 ** -9)
    gen = torch.Generator().manual_seed(3)
    deg = graph.deg.cpu()
    # the same inputs once more (the helper does not return them): rebuild g and the words from its outputs
    g = (g16_ref.float().cpu() / _scale_word(2.0 ** -9)[1])
    hid = 336
    za = torch.full((n, 2 * hid), 3.0, device="cuda")
    zb = torch.full((2 * hid + 64,), 5.0, device="cuda")
    g16 = torch.empty_like(g16_ref)
    rowmask = torch.full_like(rowmask_ref, -1)
    db = torch.zeros(256, device="cuda")
    # mask words in the layout the flag selects are not returned by the helper either: use all-zero words (dz = 0) -- the zero job
    # and g16 do not depend on them
    ntile = (n + 13) // 14
    words = torch.zeros(ntile * 256 * 4, dtype=torch.int32, device="cuda")
    fl = flag if slots8 else None
    ops._call("gnb_edge_dz_prep_wz", ops._ptr(g.cuda()), 256, ops._ptr(words), n, 256, ops._ptr(wz), ops._ptr(g16), ops._ptr(rowmask),
              ops._ptr(db), ops._ptr(fl), ops._ptr(za[:, hid:]), 2 * hid, n, hid, ops._ptr(zb), 2 * hid, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(g16, g16_ref)
    assert not db.any()
    assert bool((za[:, :hid] == 3.0).all()) and not za[:, hid:].any()
    assert not zb[: 2 * hid].any() and bool((zb[2 * hid:] == 5.0).all())
