"""CPU study for DESIGN.md section 8 item 1 (round-2 bf16 storage mode): what do outputs and gradients lose when the two
per-edge tensors that still cross HBM -- h = relu(P_i + Q_j) and dz = g_y * mask -- and the second EdgeConv weight are
kept in bf16 (as a `kind::f16` tcgen05 GEMM would need) instead of tf32-rounded fp32?

The oracle forward is run three times on the same synthetic events, weights and (fp64-forced) graphs: fp64 reference,
tf32 storage (today's training route: h, dz, W2 rounded to 10 mantissa bits) and bf16 storage (7 mantissa bits), with
fp32 arithmetic everywhere else. Errors are the per-tensor metric of the parity tests, |a - b|_inf / |b|_inf.
Lives under tests/ because it drives the oracle (test infrastructure): python tests/studies/bf16_storage_study.py
"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from graphnet_b200.synthetic import make_batch  # noqa: E402
from helpers import namespace  # noqa: E402
import oracle.dynedge_oracle as orc  # noqa: E402


def round_tf32(t):
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def round_bf16(t):
    return t.bfloat16().float()


def trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def make_edgeconv(rounder, dq_mode=None):
    """dq_mode: None = literal first Linear; "round" / "trunc" = hoisted first Linear (P = x (Wa - Wb)^T + b1, Q = x Wb^T,
    h = relu(P_i + Q_j)) with the gradient of Q rounded (today's rounding pass) or truncated (what the tensor core would do
    to an unrounded operand) to tf32 before the two GEMMs that consume it."""
    def edgeconv(x, edge_index, nn, aggr):
        lin1, _, lin2, _ = nn
        src, dst = edge_index[0], edge_index[1]
        x_i, x_j = x[dst], x[src]
        if dq_mode is None:
            h = torch.relu(lin1(torch.cat([x_i, x_j - x_i], dim=-1)))
        else:
            c = x.shape[1]
            wa, wb = lin1.weight[:, :c], lin1.weight[:, c:]
            pn = torch.nn.functional.linear(x, wa - wb, lin1.bias)
            qn = torch.nn.functional.linear(x, wb)
            if qn.requires_grad:
                pn.register_hook(round_tf32)
                qn.register_hook(round_tf32 if dq_mode == "round" else trunc_tf32)
            h = torch.relu(pn[dst] + qn[src])
        if rounder is not None:
            h = h + (rounder(h) - h).detach()                       # stored h (consumed by the edge GEMM and both gradients)
            w2 = lin2.weight + (rounder(lin2.weight) - lin2.weight).detach()
        else:
            w2 = lin2.weight
        z = torch.nn.functional.linear(h, w2, lin2.bias)
        if rounder is not None and z.requires_grad:
            z.register_hook(lambda g: rounder(g))                   # stored dz... (mask applied by relu below: g is g_y * mask)
        m = torch.relu(z)
        return m.new_zeros(x.shape[0], m.shape[1]).index_add_(0, dst, m)
    return edgeconv


def run(model, data, forced, rounder, dtype, dq_mode=None):
    saved = orc.edgeconv_ref
    orc.edgeconv_ref = make_edgeconv(rounder, dq_mode) if rounder is not None or dtype == torch.float32 else saved
    try:
        model = model.to(dtype)
        for p in model.parameters():
            p.grad = None
        d = namespace(x=data.x.to(dtype), edge_index=data.edge_index, batch=data.batch, n_pulses=data.n_pulses)
        y = model(d, forced_graphs=forced)
        y.square().sum().backward()
        return y.detach().double(), {k: p.grad.detach().double() for k, p in model.named_parameters()}
    finally:
        orc.edgeconv_ref = saved


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def main():
    raw = make_batch(24, seed=5, n_max=400)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    ptr = orc.batch_to_ptr(batch)
    ei0 = orc.knn_graph_ref(x[:, :3], 8, ptr=ptr)
    torch.manual_seed(0)
    model = orc.DynEdgeRef(7, global_pooling_schemes=["min", "max", "mean", "sum"])
    data = namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses)
    # fp64 reference + the graphs every other run is forced onto
    m64 = model.double()
    y64, inter = m64(namespace(x=x.double(), edge_index=ei0, batch=batch, n_pulses=n_pulses), return_intermediates=True)
    forced = [None] + [inter["graphs"][li] for li in range(1, 4)]
    for p in m64.parameters():
        p.grad = None
    y_ref, g_ref = run(m64, data, forced, None, torch.float64)
    print(f"events 24, pulses {x.shape[0]}, edges {ei0.shape[1]}")
    for tag, rounder, dq_mode in (("fp32 arithmetic, exact storage", None, None),
                                  ("tf32 storage of h / dz / W2", round_tf32, None),
                                  ("  + hoisted Linear, dP / dQ rounded", round_tf32, "round"),
                                  ("  + hoisted Linear, dQ TRUNCATED", round_tf32, "trunc"),
                                  ("bf16 storage of h / dz / W2", round_bf16, None)):
        y, g = run(model, data, forced, rounder, torch.float32, dq_mode)
        errs = {k: rel(g[k], g_ref[k]) for k in g_ref}
        worst = max(errs, key=errs.get)
        print(f"{tag:34s}: output {rel(y, y_ref):.2e}   gradients max {errs[worst]:.2e} ({worst})  median "
              f"{sorted(errs.values())[len(errs) // 2]:.2e}")


if __name__ == "__main__":
    main()
