"""CPU study for the fp32-grade tensor-core mode (round 2): which GEMMs of the DynEdge training step must run as
split-operand TF32 (x = x_hi + x_lo, three products) for every parameter gradient to stay within rel 1e-3 of fp64?

Every Linear of the oracle model (hoisted EdgeConv form: PQ GEMM on nodes, W2 GEMM on edges, post-processing, read-out)
is replaced by an autograd function whose three GEMMs -- forward, data gradient, weight gradient -- each take their own
operand precision:
  "1"  both operands rounded to tf32 (10 mantissa bits), fp32 accumulation       (round-1 route)
  "3"  split operands, hi*hi + hi*lo + lo*hi                                      (3xTF32)
  "x"  exact fp32
Graphs are forced from the fp64 run. Errors are the per-tensor metric of the parity tests, |a - b|_inf / |b|_inf.
Lives under tests/ because it drives the oracle (test infrastructure): python tests/studies/split_precision_study.py
"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from graphnet_b200.synthetic import make_batch  # noqa: E402
from helpers import namespace  # noqa: E402
import oracle.dynedge_oracle as orc  # noqa: E402


def rna(t):
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def trunc(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def mm(a, b, mode):
    """a [M,K] @ b [K,N] with the operand precision of `mode`."""
    if a.dtype == torch.float64 or mode == "x":
        return a @ b
    if mode == "1":
        return rna(a) @ rna(b)
    if mode == "3":
        ah, bh = trunc(a), trunc(b)
        al, bl = trunc(a - ah), trunc(b - bh)
        return ah @ bh + (ah @ bl + al @ bh)
    if mode == "1t":                                   # first operand rounded by its producer, second TRUNCATED by the
        return rna(a) @ trunc(b)                       # tensor core (an unrounded fp32 operand, e.g. h in the weight gradient)
    if mode == "3r":                                   # split with hi = rna(x) (what the kernels do)
        ah, bh = rna(a), rna(b)
        al, bl = rna(a - ah), rna(b - bh)
        return ah @ bh + (ah @ bl + al @ bh)
    if mode == "3b":                                   # hi*hi in tf32, the two correction products with bf16 operands
        ah, bh = rna(a), rna(b)
        al, bl = a - ah, b - bh
        bf = lambda t: t.bfloat16().float()
        return ah @ bh + (bf(ah) @ bf(bl) + bf(al) @ bf(bh))
    if mode == "2a":                                   # only the first operand split
        ah = trunc(a)
        al = trunc(a - ah)
        bh = rna(b)
        return ah @ bh + al @ bh
    if mode == "2b":
        bh = trunc(b)
        bl = trunc(b - bh)
        ah = rna(a)
        return ah @ bh + ah @ bl
    raise ValueError(mode)


class SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, modes):
        ctx.save_for_backward(x, w)
        ctx.modes = modes
        y = mm(x, w.t(), modes[0])
        return y if b is None else y + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        _, md, mw = ctx.modes
        gx = mm(g, w, md) if ctx.needs_input_grad[0] else None
        gw = mm(g.t(), x, mw)
        return gx, gw, g.sum(0), None


MODES = {"node": ("1", "1", "1"), "edge": ("1", "1", "1")}


def lin(x, layer, kind):
    return SplitLinear.apply(x, layer.weight, layer.bias, MODES[kind])


def edgeconv(x, edge_index, nn, aggr):
    lin1, _, lin2, _ = nn
    src, dst = edge_index[0], edge_index[1]
    c = x.shape[1]
    wa, wb = lin1.weight[:, :c], lin1.weight[:, c:]
    wcat = torch.cat([wa - wb, wb], 0)
    bcat = torch.cat([lin1.bias, torch.zeros_like(lin1.bias)])
    pq = SplitLinear.apply(x, wcat, bcat, MODES["node"])
    hid = lin1.weight.shape[0]
    h = torch.relu(pq[:, :hid][dst] + pq[:, hid:][src])
    z = lin(h, lin2, "edge")
    m = torch.relu(z)
    return m.new_zeros(x.shape[0], m.shape[1]).index_add_(0, dst, m)


class SeqWrap(torch.nn.Module):
    def __init__(self, seq):
        super().__init__()
        self.seq = seq

    def forward(self, x):
        for layer in self.seq:
            x = lin(x, layer, "node") if isinstance(layer, torch.nn.Linear) else layer(x)
        return x


def run(model, data, forced, dtype, emulate):
    saved = orc.edgeconv_ref
    post, ro = model._post_processing, model._readout
    if emulate:
        orc.edgeconv_ref = edgeconv
        model._post_processing, model._readout = SeqWrap(post), SeqWrap(ro)
    try:
        model = model.to(dtype)
        for p in model.parameters():
            p.grad = None
        d = namespace(x=data.x.to(dtype), edge_index=data.edge_index, batch=data.batch, n_pulses=data.n_pulses)
        y = model(d, forced_graphs=forced)
        y.square().sum().backward()
        return y.detach().double(), {k.replace(".seq", ""): p.grad.detach().double() for k, p in model.named_parameters()}
    finally:
        orc.edgeconv_ref = saved
        model._post_processing, model._readout = post, ro


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def main():
    nev = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    raw = make_batch(nev, seed=seed, n_max=400)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    ptr = orc.batch_to_ptr(batch)
    ei0 = orc.knn_graph_ref(x[:, :3], 8, ptr=ptr)
    torch.manual_seed(0)
    model = orc.DynEdgeRef(7, global_pooling_schemes=["min", "max", "mean", "sum"])
    data = namespace(x=x, edge_index=ei0, batch=batch, n_pulses=n_pulses)
    m64 = model.double()
    _, inter = m64(namespace(x=x.double(), edge_index=ei0, batch=batch, n_pulses=n_pulses), return_intermediates=True)
    forced = [None] + [inter["graphs"][li] for li in range(1, 4)]
    y_ref, g_ref = run(m64, data, forced, torch.float64, False)
    print(f"events {nev}, pulses {x.shape[0]}, edges {ei0.shape[1]}")
    cases = [
        ("all 1x (round-1 tf32 route)", ("1", "1", "1"), ("1", "1", "1")),
        ("fwd 3x (rna split), dgrad 1x, wgrad 1x x-truncated", ("3r", "1", "1t"), ("3r", "1", "1t")),
        ("fwd hi*hi tf32 + bf16 corrections, bwd as above", ("3b", "1", "1t"), ("3b", "1", "1t")),
    ]
    for tag, node, edge in cases:
        MODES["node"], MODES["edge"] = node, edge
        y, g = run(model, data, forced, torch.float32, True)
        errs = {k: rel(g[k], g_ref[k]) for k in g_ref}
        worst = max(errs, key=errs.get)
        print(f"{tag:54s}: out {rel(y, y_ref):.2e}  grads max {errs[worst]:.2e} ({worst})  median "
              f"{sorted(errs.values())[len(errs) // 2]:.2e}", flush=True)


if __name__ == "__main__":
    main()
