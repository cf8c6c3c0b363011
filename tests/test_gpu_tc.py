"""tcgen05 (kind::tf32) Linear kernel and the tf32 precision mode of DynEdge.

Integer-valued operands are exactly representable in tf32 and their products/sums in fp32, so the tensor
core result must be BIT-EXACT there (this pins descriptors, swizzle, TMA boxes and the epilogue mapping).
Real-valued operands: rel 2e-3 per GEMM (tf32 has 10 mantissa bits: unit round-off 2^-11 per operand)."""

import numpy as np
import pytest
import torch

from helpers import namespace, rel_err
from oracle.dynedge_oracle import DynEdgeRef, batch_to_ptr, knn_graph_ref

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tf32_mode():
    from graphnet_b200 import ops
    old = ops.PRECISION
    ops.set_precision("tf32")
    yield ops
    ops.set_precision(old)


SHAPES = [  # rows, n_out, part widths
    (128, 128, [32]), (1, 1, [4]), (300, 336, [256]), (1000, 672, [32]), (515, 256, [336]),
    (777, 336, [32, 256, 256, 256, 256]), (129, 19, [20, 7]), (4099, 128, [1024]), (64, 700, [96, 40]),
]


@pytest.mark.parametrize("rows,n_out,widths", SHAPES)
def test_tc_linear_bit_exact_on_integers(built_library, tf32_mode, rows, n_out, widths):
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
    y_ref, inter = ref(namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses), forced_graphs=forced,
                       return_intermediates=True)
    y_ref.square().sum().backward()
    errs = {f"skip{li}": rel_err(model._debug["skips"][li], inter["skips"][li]) for li in range(5)}
    errs["out"] = rel_err(y, y_ref)
    gerr = {k: rel_err(p.grad, q.grad) for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters())}
    print("tf32 rel errors:", {k: f"{v:.2e}" for k, v in errs.items()}, "max grad", f"{max(gerr.values()):.2e}")
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
