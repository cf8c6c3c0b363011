"""Precision modes "bf16" (north_star's looser, stated mode) and "bf16x3": the per-edge tensors of an EdgeConv layer
(h = relu(P_i + Q_j) and dz, the [N k, C] tensors that dominate the step's bytes) are stored as bf16 PLANES
(v ~ v0 + v1, v0 = bf16(v), v1 = bf16(v - v0)) and the three per-edge GEMMs run as tcgen05 kind::f16:
  * "bf16":   one plane, one product                      -- stated: outputs 5e-3, every gradient tensor 1.5e-2
              (the CPU emulation tests/studies/bf16_storage_study.py predicted 2.1e-3 / 5.5e-3)
  * "bf16x3": two planes, v0 w0 + v1 w0 + v0 w1           -- stated: outputs 2e-5, every gradient tensor 1e-3 (as tf32x3)
The reference computes these layers in fp32 (layers.py:55-62 -> PyG EdgeConv, dynedge.py:200-203).
Every kernel is first pinned BIT-EXACTLY on operands whose planes are exact (integers / 16-bit significands) against fp64."""

import numpy as np
import pytest
import torch

from helpers import namespace, oracle_on_kernel_decisions, rel_err, tie_heavy_events
from oracle.dynedge_oracle import DynEdgeRef, batch_to_ptr, knn_graph_ref

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ops(built_library):
    from graphnet_b200 import ops as _ops
    old = _ops.PRECISION
    yield _ops
    _ops.set_precision(old)


def planes(ops, t, np_, dst_cols=None, transpose=False):
    """bf16 planes of a CUDA fp32 matrix through gnb_to_bf16_planes."""
    t = t.contiguous()
    rows, cols = t.shape
    drows = cols if transpose else rows
    dst_cols = dst_cols or ((rows if transpose else cols) + 7) // 8 * 8
    p0 = torch.full((drows, dst_cols), 7.0, dtype=torch.bfloat16, device="cuda")
    p1 = torch.full((drows, dst_cols), 7.0, dtype=torch.bfloat16, device="cuda") if np_ == 2 else None
    ops._call("gnb_to_bf16_planes", ops._ptr(t), cols, rows, cols, ops._ptr(p0), ops._ptr(p1), dst_cols, dst_cols,
              1 if transpose else 0, ops._stream())
    return p0, p1


def two_plane_values(shape, g, scale_bits=4, mag_bits=11):
    """Values m * 2^-scale_bits with |m| < 2^mag_bits: more significand than one bf16 (8 bits), exactly two planes."""
    m = torch.randint(-(1 << mag_bits) + 1, 1 << mag_bits, shape, generator=g)
    return m.float() / float(1 << scale_bits)


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("transpose", [False, True])
def test_to_bf16_planes(ops, np_, transpose):
    g = torch.Generator().manual_seed(3)
    t = torch.randn(37, 52, generator=g)
    p0, p1 = planes(ops, t.cuda(), np_, dst_cols=64, transpose=transpose)
    src = t.t() if transpose else t
    r, c = src.shape
    want0 = src.bfloat16()
    assert torch.equal(p0.cpu()[:r, :c], want0)
    assert not p0.cpu()[:, c:].float().abs().any()
    if np_ == 2:
        assert torch.equal(p1.cpu()[:r, :c], (src - want0.float()).bfloat16())
        assert rel_err(p0.float() [:r, :c] + p1.float()[:r, :c], src) < 2.0 ** -16


def _graph(ops, sizes, seed):
    x, batch, _ = tie_heavy_events(sizes, 5, seed=seed)
    x[-15:-3] = x[-15]                                   # duplicates: degree k + 1
    ptr = batch_to_ptr(batch)
    return ops.knn_table(x.cuda(), [0, 1, 2], ptr.cuda(), 8), x.shape[0]


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("hdim", [128, 336, 40])
def test_hidden_fwd_planes_match_fp32_kernel(ops, np_, hdim):
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3], seed=hdim)
    g = torch.Generator().manual_seed(hdim)
    pq = torch.randn(n, 2 * hdim, generator=g).cuda()
    mld = 4 * ((hdim + 127) // 128)
    rows = (n + 13) // 14 * 126
    h_ref = torch.zeros(n * 9, hdim, device="cuda")
    m_ref = torch.zeros(rows, mld, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_hidden_fwd_mask", ops._ptr(pq), 2 * hdim, hdim, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n,
              ops.ACT_RELU, ops._ptr(h_ref), hdim, ops._ptr(m_ref), mld, ops._stream())
    h0 = torch.full((n * 9, hdim), 3.0, dtype=torch.bfloat16, device="cuda")
    h1 = torch.full((n * 9, hdim), 3.0, dtype=torch.bfloat16, device="cuda") if np_ == 2 else None
    m = torch.zeros(rows, mld, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_hidden_fwd_bf16", ops._ptr(pq), 2 * hdim, hdim, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n,
              ops._ptr(h0), ops._ptr(h1), hdim, ops._ptr(m), mld, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(m.cpu()[: n * 9], m_ref.cpu()[: n * 9])
    want0 = h_ref.bfloat16()
    assert torch.equal(h0, want0)
    if np_ == 2:
        assert torch.equal(h1, (h_ref - want0.float()).bfloat16())


def _agg_inputs(k, n_out, np_, which, g, n):
    """which = 'h': the activations carry two planes, weights one (and vice versa): every product of the 3-term expansion
    that is non-zero is then exact, and the fp32 sums stay below 2^24 units."""
    if np_ == 1 or which == "int":
        h = torch.randint(-1, 3, (n * 9, k), generator=g).float()
        w = torch.randint(-1, 2, (n_out, k), generator=g).float()
    elif which == "h":
        h = two_plane_values((n * 9, k), g).abs()
        w = torch.randint(-2, 3, (n_out, k), generator=g).float()
    else:
        h = torch.randint(0, 3, (n * 9, k), generator=g).float()
        w = two_plane_values((n_out, k), g)
    b = torch.randint(-3, 4, (n_out,), generator=g).float()
    return h, w, b


@pytest.mark.parametrize("np_,which", [(1, "int"), (2, "int"), (2, "h"), (2, "w")])
@pytest.mark.parametrize("k,n_out", [(336, 256), (128, 256), (40, 104), (352, 336), (64, 8)])
def test_edge_linear_agg_bf16_bit_exact(ops, np_, which, k, n_out):
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700], seed=k)
    deg = graph.deg.cpu()
    g = torch.Generator().manual_seed(n_out + k)
    h, w, b = _agg_inputs(k, n_out, np_, which, g, n)
    pre = h.double() @ w.double().t() + b.double()
    valid = (torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)).reshape(-1)
    on = (pre > 0) & valid.unsqueeze(1)
    y_ref = (pre * on).reshape(n, 9, n_out).sum(1)
    h0, h1 = planes(ops, h.cuda(), np_)
    kw = (k + 63) // 64 * 64
    w0, w1 = planes(ops, w.cuda(), np_, dst_cols=kw)
    if np_ == 2:
        assert torch.equal(h0.float().cpu() + h1.float().cpu(), h) and torch.equal((w0.float() + w1.float()).cpu()[:, :k], w)
    y = torch.empty(n, n_out, device="cuda")
    mask = torch.zeros((n + 13) // 14 * n_out * 4, dtype=torch.int32, device="cuda")
    bc = b.cuda()
    ops._call("gnb_edge_linear_agg_fwd_bf16", ops._ptr(h0), ops._ptr(h1), h0.shape[1], k, ops._ptr(w0), ops._ptr(w1), kw,
              ops._ptr(bc), ops._ptr(graph.deg), n, n_out, 0, ops._ptr(y), n_out, ops._ptr(mask), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y.cpu().double(), y_ref)
    # the mask bits drive the bf16 mask-backward kernel: dz planes of an all-ones gradient = the activation pattern
    if n_out % 8 == 0:
        gy = torch.full((n, n_out), 1.0, device="cuda")
        dz0 = torch.full((n * 9, n_out), 5.0, dtype=torch.bfloat16, device="cuda")
        dz1 = torch.full((n * 9, n_out), 5.0, dtype=torch.bfloat16, device="cuda") if np_ == 2 else None
        db = torch.zeros(n_out, device="cuda")
        ops._call("gnb_edge_mask_bwd_colsum_bf16", ops._ptr(gy), n_out, ops._ptr(mask), n, n_out, ops._ptr(dz0), ops._ptr(dz1),
                  n_out, ops._ptr(db), ops._stream())
        torch.cuda.synchronize()
        assert torch.equal(dz0.float().cpu().double(), on.double())
        assert torch.equal(db.cpu().double(), on.double().sum(0))
        if np_ == 2:
            assert not dz1.float().abs().any()


@pytest.mark.parametrize("np_", [1, 2])
def test_edge_mask_bwd_planes_match_fp32_kernel(ops, np_):
    graph, n = _graph(ops, [3, 14, 15, 200, 41], seed=1)
    n_out = 256
    g = torch.Generator().manual_seed(7)
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, ((n + 13) // 14 * n_out * 4,), generator=g, dtype=torch.int64).int().cuda()
    gy = torch.randn(n, n_out, generator=g).cuda()
    dz = torch.empty(n * 9, n_out, device="cuda")
    db_ref = torch.zeros(n_out, device="cuda")
    ops._call("gnb_edge_mask_bwd_colsum", ops._ptr(gy), n_out, ops._ptr(mask), n, n_out, ops._ptr(graph.deg), ops._ptr(dz),
              n_out, ops._ptr(db_ref), 0, ops._stream())
    dz0 = torch.empty(n * 9, n_out, dtype=torch.bfloat16, device="cuda")
    dz1 = torch.empty(n * 9, n_out, dtype=torch.bfloat16, device="cuda") if np_ == 2 else None
    db = torch.zeros(n_out, device="cuda")
    ops._call("gnb_edge_mask_bwd_colsum_bf16", ops._ptr(gy), n_out, ops._ptr(mask), n, n_out, ops._ptr(dz0), ops._ptr(dz1),
              n_out, ops._ptr(db), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dz0, dz.bfloat16())
    if np_ == 2:
        assert torch.equal(dz1, (dz - dz.bfloat16().float()).bfloat16())
    assert rel_err(db, db_ref) < 1e-5


@pytest.mark.parametrize("np_,which", [(1, "int"), (2, "int"), (2, "x"), (2, "dz")])
@pytest.mark.parametrize("rows,n_out,k_in", [(126 * 5, 256, 336), (4000, 256, 128), (77, 40, 24), (20000, 256, 336), (513, 104, 344)])
def test_wgrad_bf16_bit_exact(ops, np_, which, rows, n_out, k_in):
    g = torch.Generator().manual_seed(rows + k_in)
    if which == "x":
        x = two_plane_values((rows, k_in), g, scale_bits=4, mag_bits=9)
        dz = torch.randint(-1, 2, (rows, n_out), generator=g).float()
    elif which == "dz":
        x = torch.randint(-1, 2, (rows, k_in), generator=g).float()
        dz = two_plane_values((rows, n_out), g, scale_bits=4, mag_bits=9)
    else:
        x = torch.randint(-2, 3, (rows, k_in), generator=g).float()
        dz = torch.randint(-2, 3, (rows, n_out), generator=g).float()
    ref = dz.double().t() @ x.double()
    x0, x1 = planes(ops, x.cuda(), np_)
    z0, z1 = planes(ops, dz.cuda(), np_)
    dw = torch.zeros(n_out, k_in, device="cuda")
    ops._call("gnb_linear_bwd_weight_bf16", ops._ptr(z0), ops._ptr(z1), z0.shape[1], ops._ptr(x0), ops._ptr(x1), x0.shape[1],
              ops._ptr(dw), k_in, rows, n_out, k_in, 0, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dw.cpu().double(), ref)


@pytest.mark.parametrize("np_,which", [(1, "int"), (2, "int"), (2, "dz"), (2, "w")])
@pytest.mark.parametrize("hdim,c_out,variant", [(336, 256, 0), (128, 256, 0), (336, 256, 2), (40, 104, 0), (512, 64, 0)])
def test_dgrad_scatter_bf16_bit_exact(ops, np_, which, hdim, c_out, variant):
    """dh = dz W2, ReLU mask, dP slot sums + dQ scatter on bf16 planes, against a vectorised fp64 reference (integer-valued:
    the fp32 reductions are order independent). variant 2 forces the one-group-per-cluster kernel for two-group widths."""
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 500], seed=hdim + c_out)
    g = torch.Generator().manual_seed(hdim)
    if which == "dz":
        dz = two_plane_values((n * 9, c_out), g, scale_bits=3, mag_bits=9)
        w2 = torch.randint(-1, 2, (c_out, hdim), generator=g).float()
    elif which == "w":
        dz = torch.randint(-1, 2, (n * 9, c_out), generator=g).float()
        w2 = two_plane_values((c_out, hdim), g, scale_bits=3, mag_bits=9)
    else:
        dz = torch.randint(-1, 2, (n * 9, c_out), generator=g).float()
        w2 = torch.randint(-1, 2, (c_out, hdim), generator=g).float()
    nbr, deg = graph.nbr.cpu().long(), graph.deg.cpu()
    valid = ((torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)) & (nbr >= 0)).reshape(-1)
    dz = dz * valid.unsqueeze(1)                          # padding slots carry no gradient (the mask kernel writes zeros there)
    hbits = torch.rand(n * 9, hdim, generator=g) < 0.6
    # activation bits in the layout of gnb_edge_hidden_fwd_mask
    mld = 4 * ((hdim + 127) // 128)
    rows_m = (n + 13) // 14 * 126
    c = torch.arange(hdim)
    word, bit = 4 * (c // 128) + c % 4, (c % 128) // 4
    hm = torch.zeros(rows_m, mld, dtype=torch.int64)
    hm[: n * 9].index_put_((torch.arange(n * 9).unsqueeze(1).expand(-1, hdim), word.unsqueeze(0).expand(n * 9, -1)),
                           hbits.long() << bit.unsqueeze(0), accumulate=True)
    hm = torch.where(hm >= 2 ** 31, hm - 2 ** 32, hm).int().cuda()
    da = (dz.double() @ w2.double()) * hbits
    dp_ref = da.reshape(n, 9, hdim).sum(1)
    dq_ref = torch.zeros(n, hdim, dtype=torch.float64)
    src = nbr.reshape(-1).clamp(min=0)
    dq_ref.index_add_(0, src[valid], da[valid])
    z0, z1 = planes(ops, dz.cuda(), np_)
    cw = (c_out + 63) // 64 * 64
    wt0, wt1 = planes(ops, w2.cuda(), np_, dst_cols=cw, transpose=True)
    dq = torch.zeros(n, hdim, device="cuda")
    dp = torch.full((n, hdim), 9.0, device="cuda")
    dbias = torch.zeros(hdim, device="cuda")
    ops._call("gnb_linear_set_variant", variant)
    try:
        ops._call("gnb_edge_hidden_dgrad_scatter_bf16", ops._ptr(z0), ops._ptr(z1), z0.shape[1], c_out, ops._ptr(wt0), ops._ptr(wt1),
                  cw, ops._ptr(hm), mld, hdim, ops._ptr(graph.nbr), n, ops._ptr(dq), hdim, ops._ptr(dp), hdim, ops._ptr(dbias), 0,
                  ops._stream())
        torch.cuda.synchronize()
    finally:
        ops._call("gnb_linear_set_variant", 0)
    assert torch.equal(dp.cpu().double(), dp_ref)
    assert torch.equal(dq.cpu().double(), dq_ref)
    assert torch.equal(dbias.cpu().double(), dp_ref.sum(0))


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("rows,k,n_out", [(300, 256, 336), (4099, 1024, 128), (80000, 336, 256)])
def test_linear_fwd_bf16_grade(ops, np_, rows, k, n_out):
    """Real-valued operands against fp64: one plane = bf16 grade (8e-3), two planes = fp32 grade (2e-5)."""
    torch.manual_seed(rows)
    x, w, b = torch.randn(rows, k), torch.randn(n_out, k) / k ** 0.5, torch.randn(n_out)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    x0, x1 = planes(ops, x.cuda(), np_)
    kw = (k + 63) // 64 * 64
    w0, w1 = planes(ops, w.cuda(), np_, dst_cols=kw)
    y = torch.empty(rows, n_out, device="cuda")
    bc = b.cuda()
    ops._call("gnb_linear_fwd_bf16", ops._ptr(x0), ops._ptr(x1), x0.shape[1], k, ops._ptr(w0), ops._ptr(w1), kw, ops._ptr(bc),
              ops._ptr(y), n_out, rows, n_out, ops.ACT_RELU, 0, ops._stream())
    torch.cuda.synchronize()
    err = rel_err(y, ref)
    print(f"bf16 planes={np_} linear {rows}x{k}->{n_out}: rel {err:.2e}")
    assert err < (8e-3 if np_ == 1 else 2e-5)


def _scale_word(maxval):
    """Device word with the fp32 bits of `maxval` and the power-of-two scale gnb_pow2_scale derives from it."""
    bits = int(np.float32(maxval).view(np.uint32))
    e = (bits >> 23) & 0xFF
    scale = 2.0 ** (14 - (e - 127))
    return torch.tensor([bits], dtype=torch.int64).int().cuda(), scale


def test_absmax_bits(ops):
    g = torch.Generator().manual_seed(11)
    a = torch.randn(1000, 256, generator=g) * 3e-4
    a[977, 13] = -0.0421
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    ac = a.cuda()
    ops._call("gnb_absmax_bits", ops._ptr(ac), 256, 1000, 256, 0, ops._ptr(word), ops._stream())
    torch.cuda.synchronize()
    assert int(word.item()) == int(np.float32(0.0421).view(np.uint32))
    ops._call("gnb_absmax_bits", ops._ptr(ac), 256, 1000, 256, 1, ops._ptr(word), ops._stream())      # max with 2 x the maximum
    torch.cuda.synchronize()
    assert int(word.item()) == int(np.float32(0.0842).view(np.uint32))


@pytest.mark.parametrize("maxval", [0.0421, 3.7e-6, 900.0])
def test_edge_mask_bwd_f16_is_the_scaled_fp32_kernel(ops, maxval):
    graph, n = _graph(ops, [3, 14, 15, 200, 41], seed=1)
    n_out = 256
    g = torch.Generator().manual_seed(7)
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, ((n + 13) // 14 * n_out * 4,), generator=g, dtype=torch.int64).int().cuda()
    gy = (torch.rand(n, n_out, generator=g) * 2 - 1) * maxval
    gy[0, 0] = maxval
    gy = gy.cuda()
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops._call("gnb_absmax_bits", ops._ptr(gy), n_out, n, n_out, 0, ops._ptr(word), ops._stream())
    _, scale = _scale_word(maxval)
    assert 2 ** 14 <= maxval * scale < 2 ** 15
    dz = torch.empty(n * 9, n_out, device="cuda")
    db_ref = torch.zeros(n_out, device="cuda")
    ops._call("gnb_edge_mask_bwd_colsum", ops._ptr(gy), n_out, ops._ptr(mask), n, n_out, ops._ptr(graph.deg), ops._ptr(dz),
              n_out, ops._ptr(db_ref), 0, ops._stream())
    dz16 = torch.empty(n * 9, n_out, dtype=torch.float16, device="cuda")
    db = torch.zeros(n_out, device="cuda")
    ops._call("gnb_edge_mask_bwd_colsum_f16", ops._ptr(gy), n_out, ops._ptr(mask), n, n_out, ops._ptr(dz16), n_out, ops._ptr(db),
              ops._ptr(word), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dz16, (dz * scale).half())
    assert rel_err(dz16.float() / scale, dz) < 2.0 ** -11          # fp16 = tf32's significand
    assert rel_err(db, db_ref) < 1e-5


def f16_planes(ops, t, np_, dst_cols=None, transpose=False):
    t = t.contiguous()
    rows, cols = t.shape
    drows = cols if transpose else rows
    dst_cols = dst_cols or ((rows if transpose else cols) + 7) // 8 * 8
    p0 = torch.full((drows, dst_cols), 7.0, dtype=torch.float16, device="cuda")
    p1 = torch.full((drows, dst_cols), 7.0, dtype=torch.float16, device="cuda") if np_ == 2 else None
    ops._call("gnb_to_f16_planes", ops._ptr(t), cols, rows, cols, ops._ptr(p0), ops._ptr(p1), dst_cols, dst_cols,
              1 if transpose else 0, ops._stream())
    return p0, p1


@pytest.mark.parametrize("np_", [1, 2])
def test_to_f16_planes(ops, np_):
    g = torch.Generator().manual_seed(3)
    t = torch.randn(37, 52, generator=g) * 0.05
    p0, p1 = f16_planes(ops, t.cuda(), np_, dst_cols=64)
    assert torch.equal(p0.cpu()[:, :52], t.half()) and not p0.cpu()[:, 52:].float().abs().any()
    if np_ == 2:
        assert torch.equal(p1.cpu()[:, :52], (t - t.half().float()).half())
        assert rel_err(p0.float()[:, :52] + p1.float()[:, :52], t) < 2.0 ** -19


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("hdim", [128, 336])
def test_hidden_fwd_f16_planes_are_the_scaled_fp32_kernel(ops, np_, hdim):
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3], seed=hdim)
    g = torch.Generator().manual_seed(hdim)
    pq = (torch.randn(n, 2 * hdim, generator=g) * 37.0).cuda()
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops._call("gnb_absmax_bits", ops._ptr(pq), 2 * hdim, n, 2 * hdim, 1, ops._ptr(word), ops._stream())
    bound = 2 * float(pq.abs().max())
    _, scale = _scale_word(bound)
    mld = 4 * ((hdim + 127) // 128)
    rows = (n + 13) // 14 * 126
    h_ref = torch.zeros(n * 9, hdim, device="cuda")
    m_ref = torch.zeros(rows, mld, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_hidden_fwd_mask", ops._ptr(pq), 2 * hdim, hdim, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n,
              ops.ACT_RELU, ops._ptr(h_ref), hdim, ops._ptr(m_ref), mld, ops._stream())
    h0 = torch.full((n * 9, hdim), 3.0, dtype=torch.float16, device="cuda")
    h1 = torch.full((n * 9, hdim), 3.0, dtype=torch.float16, device="cuda") if np_ == 2 else None
    m = torch.zeros(rows, mld, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_hidden_fwd_f16", ops._ptr(pq), 2 * hdim, hdim, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n,
              ops._ptr(h0), ops._ptr(h1), hdim, ops._ptr(m), mld, ops._ptr(word), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(m.cpu()[: n * 9], m_ref.cpu()[: n * 9])
    assert float(h_ref.max()) * scale < 2 ** 15 and not torch.isinf(h0.float()).any()
    want0 = (h_ref * scale).half()
    assert torch.equal(h0, want0)
    if np_ == 2:
        assert torch.equal(h1, (h_ref * scale - want0.float()).half())
        assert rel_err((h0.float() + h1.float()) / scale, h_ref) < 2.0 ** -20


@pytest.mark.parametrize("np_,which", [(1, "int"), (2, "int"), (2, "h"), (2, "w")])
@pytest.mark.parametrize("k,n_out", [(336, 256), (128, 256), (40, 104)])
def test_edge_linear_agg_f16_bit_exact(ops, np_, which, k, n_out):
    """fp16 planes with a power-of-two scale on h: same exact cases as the bf16 planes (fp16 holds 11-bit significands, so the
    two-plane values fit as well), the epilogue's 2^-s is exact."""
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700], seed=k)
    deg = graph.deg.cpu()
    g = torch.Generator().manual_seed(n_out + k)
    h, w, b = _agg_inputs(k, n_out, np_, which, g, n)
    if which == "h":
        h = two_plane_values((n * 9, k), g, scale_bits=4, mag_bits=15).abs()        # 15-bit significands: two fp16 planes
    pre = h.double() @ w.double().t() + b.double()
    valid = (torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)).reshape(-1)
    on = (pre > 0) & valid.unsqueeze(1)
    y_ref = (pre * on).reshape(n, 9, n_out).sum(1)
    word, scale = _scale_word(float(h.max()))
    h0, h1 = f16_planes(ops, (h * scale).cuda(), np_)
    kw = (k + 63) // 64 * 64
    w0, w1 = f16_planes(ops, w.cuda(), np_, dst_cols=kw)
    if np_ == 2:
        assert torch.equal((h0.float().cpu() + h1.float().cpu()) / scale, h) and torch.equal((w0.float() + w1.float()).cpu()[:, :k], w)
    y = torch.empty(n, n_out, device="cuda")
    mask = torch.zeros((n + 13) // 14 * n_out * 4, dtype=torch.int32, device="cuda")
    bc = b.cuda()
    ops._call("gnb_edge_linear_agg_fwd_f16", ops._ptr(h0), ops._ptr(h1), h0.shape[1], k, ops._ptr(w0), ops._ptr(w1), kw,
              ops._ptr(bc), ops._ptr(graph.deg), n, n_out, 0, ops._ptr(y), n_out, ops._ptr(mask), ops._ptr(word), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y.cpu().double(), y_ref)


@pytest.mark.parametrize("planes_x", [1, 2])
@pytest.mark.parametrize("rows,n_out,k_in", [(126 * 5, 256, 336), (4000, 256, 128), (20000, 256, 336), (513, 104, 344)])
def test_wgrad_f16_bit_exact(ops, planes_x, rows, n_out, k_in):
    """x as scaled fp16 plane(s), dz as one scaled fp16 plane: the epilogue's two inverse powers of two are exact."""
    g = torch.Generator().manual_seed(rows + k_in)
    x = two_plane_values((rows, k_in), g, scale_bits=4, mag_bits=13) if planes_x == 2 else torch.randint(-2, 3, (rows, k_in), generator=g).float()
    dz = torch.randint(-3, 4, (rows, n_out), generator=g).float() * 2.0 ** -9
    if planes_x == 2:
        dz = torch.randint(-1, 2, (rows, n_out), generator=g).float() * 2.0 ** -9
    wz, sz = _scale_word(3 * 2.0 ** -9)
    wx, sx = _scale_word(float(x.abs().max()))
    ref = dz.double().t() @ x.double()
    x0, x1 = f16_planes(ops, (x * sx).cuda(), planes_x)
    dz16 = (dz * sz).half().cuda()
    dw = torch.zeros(n_out, k_in, device="cuda")
    ops._call("gnb_linear_bwd_weight_f16", ops._ptr(dz16), n_out, ops._ptr(x0), ops._ptr(x1), x0.shape[1], ops._ptr(dw), k_in,
              rows, n_out, k_in, ops._ptr(wz), ops._ptr(wx), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dw.cpu().double(), ref)


@pytest.mark.parametrize("hdim,c_out,variant", [(336, 256, 0), (128, 256, 0), (336, 256, 2), (40, 104, 0)])
def test_dgrad_scatter_f16_bit_exact(ops, hdim, c_out, variant):
    graph, n = _graph(ops, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 500], seed=hdim + c_out)
    g = torch.Generator().manual_seed(hdim)
    dz = torch.randint(-1, 2, (n * 9, c_out), generator=g).float() * 2.0 ** -20
    w2 = torch.randint(-8, 9, (c_out, hdim), generator=g).float() / 8          # exact in fp16
    word, scale = _scale_word(2.0 ** -20)
    nbr, deg = graph.nbr.cpu().long(), graph.deg.cpu()
    valid = ((torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)) & (nbr >= 0)).reshape(-1)
    dz = dz * valid.unsqueeze(1)
    hbits = torch.rand(n * 9, hdim, generator=g) < 0.6
    mld = 4 * ((hdim + 127) // 128)
    rows_m = (n + 13) // 14 * 126
    c = torch.arange(hdim)
    word_i, bit = 4 * (c // 128) + c % 4, (c % 128) // 4
    hm = torch.zeros(rows_m, mld, dtype=torch.int64)
    hm[: n * 9].index_put_((torch.arange(n * 9).unsqueeze(1).expand(-1, hdim), word_i.unsqueeze(0).expand(n * 9, -1)),
                           hbits.long() << bit.unsqueeze(0), accumulate=True)
    hm = torch.where(hm >= 2 ** 31, hm - 2 ** 32, hm).int().cuda()
    da = (dz.double() @ w2.double()) * hbits
    dp_ref = da.reshape(n, 9, hdim).sum(1)
    dq_ref = torch.zeros(n, hdim, dtype=torch.float64)
    dq_ref.index_add_(0, nbr.reshape(-1).clamp(min=0)[valid], da[valid])
    dz16 = (dz * scale).half().cuda()
    cw = (c_out + 63) // 64 * 64
    wt16, _ = f16_planes(ops, w2.cuda(), 1, dst_cols=cw, transpose=True)
    assert torch.equal(wt16.cpu()[:, :c_out].float(), w2.t()) and not wt16.cpu()[:, c_out:].float().abs().any()
    dq = torch.zeros(n, hdim, device="cuda")
    dp = torch.full((n, hdim), 9.0, device="cuda")
    dbias = torch.zeros(hdim, device="cuda")
    ops._call("gnb_linear_set_variant", variant)
    try:
        ops._call("gnb_edge_hidden_dgrad_scatter_f16", ops._ptr(dz16), c_out, c_out, ops._ptr(wt16), cw, ops._ptr(hm), mld, hdim,
                  ops._ptr(graph.nbr), n, ops._ptr(dq), hdim, ops._ptr(dp), hdim, ops._ptr(dbias), 0, ops._ptr(word), ops._stream())
        torch.cuda.synchronize()
    finally:
        ops._call("gnb_linear_set_variant", 0)
    assert torch.equal(dp.cpu().double(), dp_ref)
    assert torch.equal(dq.cpu().double(), dq_ref)
    assert torch.equal(dbias.cpu().double(), dp_ref.sum(0))


def _mask_case(ops, sizes, n_out, seed, gmax):
    """Random ReLU bit masks in the aggregating epilogue's layout (bits of padding slots zero), a gradient g, and the dz they
    imply: dz[(i, s), c] = g[i, c] * bit."""
    graph, n = _graph(ops, sizes, seed=seed)
    gen = torch.Generator().manual_seed(seed)
    deg = graph.deg.cpu()
    valid = (torch.arange(9).unsqueeze(0) < deg.unsqueeze(1))                       # [n, 9]
    bits = (torch.rand(n, 9, n_out, generator=gen) < 0.55) & valid.unsqueeze(2)     # [n, 9, C]
    ntile = (n + 13) // 14
    full = torch.zeros(ntile * 14, 9, n_out, dtype=torch.bool)
    full[:n] = bits
    cols = full.reshape(ntile, 126, n_out)                                           # bit index 9 f + s
    words = torch.zeros(ntile, n_out, 4, dtype=torch.int64)
    for w in range(4):
        chunk = cols[:, 32 * w: min(126, 32 * w + 32)]                               # [ntile, <=32, C]
        sh = torch.arange(chunk.shape[1]).view(1, -1, 1)
        words[:, :, w] = (chunk.long() << sh).sum(1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).int().reshape(-1).cuda()
    # the row-major copy: word c / 32 of row (i, s), bit c % 32 (rows padded to whole tiles)
    rowm = None
    if n_out % 32 == 0:
        rb = full.reshape(ntile * 126, n_out // 32, 32).long()
        rowm = (rb << torch.arange(32).view(1, 1, 32)).sum(2)
        rowm = torch.where(rowm >= 2 ** 31, rowm - 2 ** 32, rowm).int().cuda()
    g = torch.randint(-4, 5, (n, n_out), generator=gen).float() * (gmax / 4)
    dz = (g.unsqueeze(1) * bits).reshape(n * 9, n_out)
    # the kernels' inputs come from gnb_edge_dz_prep: fp16(g 2^s), row-major bits, bias gradient
    wz, sz = _scale_word(gmax)
    gc = g.cuda()
    g16 = torch.full((n, n_out), 7.0, dtype=torch.float16, device="cuda")
    rowmask = torch.full((ntile * 126, max(n_out // 32, 1)), -1, dtype=torch.int32, device="cuda")
    db = torch.zeros(n_out, device="cuda")
    if n_out % 32 == 0:
        ops._call("gnb_edge_dz_prep", ops._ptr(gc), n_out, ops._ptr(words), n, n_out, ops._ptr(wz), ops._ptr(g16), ops._ptr(rowmask),
                  ops._ptr(db), ops._stream())
        torch.cuda.synchronize()
        assert torch.equal(g16.cpu(), (g * sz).half())
        assert torch.equal(rowmask.cpu(), rowm.cpu())
        assert torch.equal(db.cpu().double(), dz.double().sum(0))
    return graph, n, wz, g16, dz, rowmask


@pytest.mark.parametrize("n_out,k_in,sizes", [(256, 336, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700]), (256, 128, [400, 900, 14, 15]),
                                              (96, 344, [77, 5, 230]), (256, 336, [3000, 2500])])
def test_wgrad_f16_masked_bit_exact(ops, n_out, k_in, sizes):
    """dW2 with dz expanded inside the kernel from g and the mask words == the same product on a stored dz (fp64 reference)."""
    graph, n, wz, g16, dz, rowm = _mask_case(ops, sizes, n_out, seed=n_out + k_in, gmax=2.0 ** -9)
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(-2, 3, (n * 9, k_in), generator=gen).float()
    wx, sx = _scale_word(2.0)
    ref = dz.double().t() @ x.double()
    x16 = (x * sx).half().cuda()
    dw = torch.zeros(n_out, k_in, device="cuda")
    ops._call("gnb_linear_bwd_weight_f16_masked", ops._ptr(g16), ops._ptr(rowm), ops._ptr(x16), k_in, ops._ptr(dw), k_in, n,
              n_out, k_in, ops._ptr(wz), ops._ptr(wx), ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dw.cpu().double(), ref)


@pytest.mark.parametrize("hdim,c_out,sizes", [(336, 256, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 500]), (128, 256, [400, 900, 14, 15]),
                                              (40, 128, [77, 5, 230]), (336, 256, [3000, 2500]), (512, 64, [100, 37])])
def test_dgrad_scatter_f16_masked_bit_exact(ops, hdim, c_out, sizes):
    graph, n, word, g16, dz, rowm = _mask_case(ops, sizes, c_out, seed=hdim + c_out, gmax=2.0 ** -20)
    gen = torch.Generator().manual_seed(hdim)
    w2 = torch.randint(-8, 9, (c_out, hdim), generator=gen).float() / 8
    nbr, deg = graph.nbr.cpu().long(), graph.deg.cpu()
    valid = ((torch.arange(9).unsqueeze(0) < deg.unsqueeze(1)) & (nbr >= 0)).reshape(-1)
    hbits = torch.rand(n * 9, hdim, generator=gen) < 0.6
    mld = 4 * ((hdim + 127) // 128)
    rows_m = (n + 13) // 14 * 126
    c = torch.arange(hdim)
    word_i, bit = 4 * (c // 128) + c % 4, (c % 128) // 4
    hm = torch.zeros(rows_m, mld, dtype=torch.int64)
    hm[: n * 9].index_put_((torch.arange(n * 9).unsqueeze(1).expand(-1, hdim), word_i.unsqueeze(0).expand(n * 9, -1)),
                           hbits.long() << bit.unsqueeze(0), accumulate=True)
    hm = torch.where(hm >= 2 ** 31, hm - 2 ** 32, hm).int().cuda()
    da = (dz.double() @ w2.double()) * hbits
    dp_ref = da.reshape(n, 9, hdim).sum(1)
    dq_ref = torch.zeros(n, hdim, dtype=torch.float64)
    dq_ref.index_add_(0, nbr.reshape(-1).clamp(min=0)[valid], da[valid])
    cw = (c_out + 63) // 64 * 64
    wt16, _ = f16_planes(ops, w2.cuda(), 1, dst_cols=cw, transpose=True)
    dq = torch.zeros(n, hdim, device="cuda")
    dp = torch.full((n, hdim), 9.0, device="cuda")
    dbias = torch.zeros(hdim, device="cuda")
    ops._call("gnb_edge_hidden_dgrad_scatter_f16_masked", ops._ptr(g16), ops._ptr(rowm), c_out, ops._ptr(wt16), cw, ops._ptr(hm),
              mld, hdim, ops._ptr(graph.nbr), n, ops._ptr(dq), hdim, ops._ptr(dp), hdim, ops._ptr(dbias), 0, ops._ptr(word),
              ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(dp.cpu().double(), dp_ref)
    assert torch.equal(dq.cpu().double(), dq_ref)
    assert torch.equal(dbias.cpu().double(), dp_ref.sum(0))


def _pq_lane_interleaved(pq, hid):
    """pq_layout 1 of gnb_edgeconv_fused_fwd_f16: inside every full 64-column block of the P and of the Q half, stored column s
    holds hidden unit 8 (s / 4) + s % 4 (s < 32) or 8 ((s - 32) / 4) + 4 + s % 4 (include/graphnet_b200.h)."""
    idx = torch.arange(2 * hid)
    s = torch.arange(64)
    unit = torch.where(s < 32, 8 * (s // 4) + s % 4, 8 * ((s - 32) // 4) + 4 + s % 4)
    for half in (0, hid):
        for blk in range(0, hid - 63, 64):
            idx[half + blk: half + blk + 64] = half + blk + unit
    return pq[:, idx].contiguous()


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("hid,n_out,sizes", [(336, 256, [1, 2, 5, 9, 10, 64, 130, 12, 300, 3, 700]), (128, 256, [400, 900, 14, 15]),
                                             (40, 104, [77, 5, 230]), (336, 256, [3000, 2500]), (128, 256, [14 * 5])])
def test_edgeconv_fused_fwd_f16_matches_the_two_kernel_forward(ops, np_, hid, n_out, sizes, layout):
    """gather + hidden layer + second Linear + aggregation in one kernel == hidden-layer kernel followed by the aggregating GEMM:
    the two run the same products on the same fp16 planes, so y agrees to fp32 summation order of the TMEM accumulators (exact
    on these integer-valued operands) and plane 0 / the ReLU bits / the mask words agree bit for bit."""
    graph, n = _graph(ops, sizes, seed=hid + n_out)
    gen = torch.Generator().manual_seed(hid)
    pq = torch.randint(-6, 7, (n, 2 * hid), generator=gen).float() * 0.25
    w = (torch.randint(-2, 3, (n_out, hid), generator=gen).float() if np_ == 1 else two_plane_values((n_out, hid), gen, scale_bits=6, mag_bits=13))
    b = torch.randint(-3, 4, (n_out,), generator=gen).float()
    pqc, bc = pq.cuda(), b.cuda()
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops._call("gnb_absmax_bits", ops._ptr(pqc), 2 * hid, n, 2 * hid, 1, ops._ptr(word), ops._stream())
    kw = (hid + 63) // 64 * 64
    w0, w1 = f16_planes(ops, w.cuda(), np_, dst_cols=kw)
    ntile = (n + 13) // 14
    mld = 4 * ((hid + 127) // 128)
    # reference: the two-kernel forward
    h0 = torch.zeros(n * 9, hid, dtype=torch.float16, device="cuda")
    h1 = torch.zeros(n * 9, hid, dtype=torch.float16, device="cuda") if np_ == 2 else None
    hm = torch.zeros(ntile * 126, mld, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_hidden_fwd_f16", ops._ptr(pqc), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), 9, n, ops._ptr(h0),
              ops._ptr(h1), hid, ops._ptr(hm), mld, ops._ptr(word), ops._stream())
    y_ref = torch.empty(n, n_out, device="cuda")
    m_ref = torch.zeros(ntile * n_out * 4, dtype=torch.int32, device="cuda")
    ops._call("gnb_edge_linear_agg_fwd_f16", ops._ptr(h0), ops._ptr(h1), hid, hid, ops._ptr(w0), ops._ptr(w1), kw, ops._ptr(bc),
              ops._ptr(graph.deg), n, n_out, 0, ops._ptr(y_ref), n_out, ops._ptr(m_ref), ops._ptr(word), ops._stream())
    # fused (layout 1: the same PQ with its columns stored lane-interleaved; every output stays in natural order)
    if layout == 1:
        pqc = _pq_lane_interleaved(pq, hid).cuda()
    y = torch.empty(n, n_out, device="cuda")
    m = torch.zeros(ntile * n_out * 4, dtype=torch.int32, device="cuda")
    h0f = torch.full((n * 9, hid), 5.0, dtype=torch.float16, device="cuda")
    guard = torch.full((4096,), 77, dtype=torch.uint8, device="cuda")          # right behind the side outputs: must stay untouched
    hb_all = torch.zeros(ntile * 126 * mld * 4 + 4096, dtype=torch.uint8, device="cuda")
    hb_all[ntile * 126 * mld * 4:] = 77
    hb = hb_all[: ntile * 126 * mld * 4].view(ntile * 126, mld * 4)
    ops._call("gnb_edgeconv_fused_fwd_f16", ops._ptr(pqc), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), n, ops._ptr(w0),
              ops._ptr(w1), kw, ops._ptr(bc), n_out, 0, ops._ptr(y), n_out, ops._ptr(m), ops._ptr(h0f), hid, ops._ptr(hb), mld * 4,
              ops._ptr(word), layout, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(hb_all[ntile * 126 * mld * 4:], guard)
    assert torch.equal(h0f, h0)
    assert torch.equal(y, y_ref)
    assert torch.equal(m, m_ref)
    # activation bits: the fused kernel's row-major bytes (bit c % 8 of byte c / 8) vs h > 0
    bits = ((hb.cpu()[: n * 9, : hid // 8].long().unsqueeze(2) >> torch.arange(8).view(1, 1, 8)) & 1).reshape(n * 9, hid).bool()
    assert torch.equal(bits, h0.cpu().float() > 0) or torch.equal(bits, (h0.cpu().float() > 0) | bits)   # (fp16 may flush a tiny h to 0)
    # inference form: no side outputs
    y2 = torch.empty(n, n_out, device="cuda")
    ops._call("gnb_edgeconv_fused_fwd_f16", ops._ptr(pqc), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), n, ops._ptr(w0),
              ops._ptr(w1), kw, ops._ptr(bc), n_out, 0, ops._ptr(y2), n_out, ops._ptr(None), ops._ptr(None), hid, ops._ptr(None), mld * 4,
              ops._ptr(word), layout, ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(y2, y_ref)


def test_mixed16_unfused_forward_matches_the_fused_forward(ops):
    """Executor flag bit 3 (ops.UNFUSED_FORWARD): two-kernel forward vs fused EdgeConv forward, training step gradients."""
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    ops.set_precision("mixed16")
    raw = make_batch(24, seed=11, n_max=500)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    torch.manual_seed(3)
    model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
    data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
    model._debug_record = True
    outs, first, grads = [], [], []
    for unfused in (False, True):
        ops.UNFUSED_FORWARD = unfused
        try:
            model.zero_grad(set_to_none=True)
            y = model(data)
            first.append(model._debug["skips"][1].detach().clone())
            y.square().sum().backward()
            outs.append(y.detach().clone())
            grads.append([p.grad.clone() for p in model.parameters()])
        finally:
            ops.UNFUSED_FORWARD = False
    # the first layer runs on the same graph in both: same products, TMEM summation order aside. Later layers rebuild their kNN
    # graphs from features that differ in the last bits, so a near-tie may pick another neighbour: compared at the mode's tolerance
    assert rel_err(first[0], first[1]) < 2e-6
    assert rel_err(outs[0], outs[1]) < 1e-3
    for a, b in zip(*grads):
        assert rel_err(a, b) < 2e-2


def test_mixed16_stored_dz_route_matches_the_masked_route(ops):
    """Executor flag bit 2 (ops.STORE_DZ): the stored-dz backward and the in-kernel expansion compute the same gradients."""
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    ops.set_precision("mixed16")
    raw = make_batch(24, seed=11, n_max=500)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    torch.manual_seed(3)
    model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
    data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
    grads = []
    # both runs on the two-kernel forward: the stored-dz route cannot use the fused forward kernel (the executor fuses the
    # forward only where the backward expands dz itself), and a different forward kernel means a different TMEM summation order
    ops.UNFUSED_FORWARD = True
    try:
        for store in (False, True):
            ops.STORE_DZ = store
            model.zero_grad(set_to_none=True)
            model(data).square().sum().backward()
            grads.append([p.grad.clone() for p in model.parameters()])
    finally:
        ops.STORE_DZ = False
        ops.UNFUSED_FORWARD = False
    for a, b in zip(*grads):
        assert rel_err(a, b) < 2e-5


MODE_TOL = {"bf16": (5e-3, 1.5e-2), "bf16x3": (2e-5, 1e-3), "mixed16": (2e-5, 1e-3), "f16": (1e-3, 3e-3)}


@pytest.mark.parametrize("mode", ["bf16", "bf16x3", "mixed16", "f16"])
def test_dynedge_bf16_modes_vs_oracle(ops, mode):
    """Default DynEdge (4 pooling schemes) on the executor route against the fp64 oracle fed the kernel's own graphs and
    read-out decisions: outputs and EVERY parameter gradient within the mode's stated tolerance; latent kNN graphs bit-exact
    on the kernel's own features."""
    ops.set_precision(mode)
    assert ops.USE_EXECUTOR
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(48 if mode != "f16" else 96, seed=5, n_max=400)      # f16 = tf32 grade: held on >= 2 778 pulses like tf32
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
    ref = ref.double()
    y_ref, inter, _ = oracle_on_kernel_decisions(ref, namespace(x=x.double(), edge_index=ei0, batch=batch, n_pulses=n_pulses),
                                                 forced, y, mode)
    y_ref.square().sum().backward()
    errs = {f"skip{li}": rel_err(model._debug["skips"][li], inter["skips"][li]) for li in range(5)}
    errs["out"] = rel_err(y, y_ref)
    gerr = {k: rel_err(p.grad, q.grad) for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters())}
    print(f"{mode} rel errors:", {k: f"{v:.2e}" for k, v in errs.items()}, "max grad", f"{max(gerr.values()):.2e}")
    print(f"{mode} grad rel errors:", {k: f"{v:.1e}" for k, v in gerr.items()})
    out_tol, grad_tol = MODE_TOL[mode]
    assert errs["out"] < out_tol, errs
    assert max(gerr.values()) < grad_tol, gerr


@pytest.mark.parametrize("mode", ["bf16", "bf16x3", "mixed16", "f16"])
def test_dynedge_bf16_inference_matches_training_forward(ops, mode):
    """The inference route (shared per-edge buffers, no masks) computes the same numbers as the training forward."""
    ops.set_precision(mode)
    from graphnet_b200 import Data
    from graphnet_b200.models.gnn import DynEdge
    from graphnet_b200.models.graphs.edges import KNNEdges
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(16, seed=9, n_max=300)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    torch.manual_seed(1)
    model = DynEdge(7, global_pooling_schemes=["min", "max", "mean", "sum"]).cuda()
    data = KNNEdges(8)(Data(x=x.cuda(), batch=batch.cuda(), n_pulses=n_pulses.cuda()))
    y_train = model(data)
    with torch.no_grad():
        y_inf = model(data)
    if mode == "f16":
        # f16 trains on the fused forward (8-slot tiles where the graph allows) and infers on the two-kernel forward (9-slot
        # tiles): same products, but a node's eight messages are summed in a tile-position dependent order
        assert rel_err(y_train, y_inf) < 1e-3
    else:
        assert torch.equal(y_train.detach(), y_inf)
