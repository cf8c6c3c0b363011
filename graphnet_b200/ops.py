"""Torch-facing operators of the B200 DynEdge path.

Every function here hands raw device pointers, sizes and the current CUDA
stream to the C-ABI library (`include/graphnet_b200.h`); PyTorch only owns the
memory and the autograd graph. Nothing in this module computes on the CPU: a
non-CUDA tensor raises.
"""

from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2        # LEAKY: torch.nn.LeakyReLU() default slope 0.01
FLAG_ROUND_TF32 = 0x100
AGGR = {"add": 0, "sum": 0, "mean": 1, "max": 2}
POOL = {"min": 0, "max": 1, "sum": 2, "mean": 3}

# Precision of the dense contractions:
#   "fp32": fp32 SIMT GEMMs (exact fp32 products, ~1e-6 from the oracle)
#   "tf32": tcgen05 kind::tf32 GEMMs (operands rounded to tf32 with cvt.rna, fp32 accumulation in TMEM)
#   "tf32x3": fp32-grade tensor-core mode -- every FORWARD GEMM runs on split operands (x = x_hi + x_lo, W = W_hi + W_lo,
#             three kind::tf32 products per K step, activations stay plain fp32); the backward GEMMs are single-pass tf32.
#             Outputs ~1e-6, gradients ~5e-4 from the fp64 oracle (tests/studies/split_precision_study.py: the forward's
#             rounding, not the backward's, is what moved the tf32 gradients to 2e-3).
#   "bf16": north_star's looser mode -- the per-edge tensors of the executor route (h, dz: the [N k, C] tensors that dominate
#             the bytes) are stored as ONE bf16 plane and the three per-edge GEMMs run as kind::f16; node-level GEMMs as "tf32".
#             Stated tolerance: outputs 5e-3, gradients 1.5e-2 (tests/test_gpu_bf16.py).
#   "bf16x3": the per-edge tensors as TWO bf16 planes (v ~ bf16(v) + bf16(v - bf16(v))), three products per per-edge GEMM in the
#             forward AND the backward pass; node-level forward GEMMs as "tf32x3". fp32 grade like "tf32x3".
#             Both bf16 modes exist on the executor route (k = 8 graphs); elsewhere they behave like "tf32" / "tf32x3".
PRECISION = os.environ.get("GNB_PRECISION", "fp32")
#   "mixed16": the per-edge tensors as fp16 planes scaled per layer by a power of two (fp16 = tf32's 11-bit significand in half
#             the bytes; the scale word comes from an absmax pass and is undone exactly in the consuming epilogues): forward on TWO
#             planes of h and W2 (three products, fp32 grade), backward on ONE plane of dz, h and W2^T (tf32 grade); node-level
#             forward GEMMs as "tf32x3". Same grade as "tf32x3" (fp32-grade forward, tf32-grade backward). bench.py's headline.
#   "f16": ONE scaled fp16 plane per per-edge tensor, forward and backward: the grade of "tf32" (same 11-bit significand) at half
#             the per-edge bytes and twice the MMA rate; node-level GEMMs as "tf32". The inference mode of bench.py.
PRECISIONS = ("fp32", "tf32", "tf32x3", "bf16", "bf16x3", "mixed16", "f16")


def set_precision(mode: str) -> None:
    global PRECISION
    if mode not in PRECISIONS:
        raise ValueError(f"unknown precision {mode!r}")
    PRECISION = mode


def _tf32() -> bool:
    """Tensor-core GEMMs (every mode but fp32)."""
    return PRECISION != "fp32"


def _split() -> bool:
    return PRECISION in ("tf32x3", "bf16x3", "mixed16")


def _fround() -> bool:
    """Forward activations that feed a tensor-core GEMM are stored rounded to tf32 (single-pass mode only)."""
    return PRECISION in ("tf32", "bf16", "f16")


def _mark_rounded(t: Tensor) -> Tensor:
    t._gnb_tf32 = True          # values are exactly representable in tf32
    return t


def _is_rounded(t: Tensor) -> bool:
    return bool(getattr(t, "_gnb_tf32", False))

# fp16-plane modes: True keeps the stored-dz backward (mask-backward kernel writes dz [N k, C]); False (default) expands dz
# inside the weight-gradient and scattering GEMMs from the node-level gradient and the ReLU bit mask
STORE_DZ = False
# fp16-plane modes: False (default) runs the fused EdgeConv forward (gather + hidden layer + second Linear + aggregation in one
# kernel, h built in shared memory); True keeps the two-kernel forward (hidden-layer kernel writes h, aggregating GEMM reads it)
UNFUSED_FORWARD = False
# fp16-plane modes: False (default) lets the device pick the 8-slot edge layout for every graph without a k + 1-neighbour node
# (16 nodes x 8 slots per tile, -11 % rows in every per-edge kernel); True always uses the 9-slot layout
SLOTS9 = False

# number of kernels launched through this module (bench.py reports it as `gpu_launches`)
LAUNCHES = 0


def _call(name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += 1
    _lib.check(getattr(_lib.load(), name)(*args), name)


def kernel_launch_count() -> int:
    """Kernels launched by libgraphnet_b200.so so far (counted inside the library, executor included)."""
    return int(_lib.load().gnb_launch_count())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("graphnet_b200 operators run on CUDA tensors only (no CPU fallback)")


def _rowmajor(t: Tensor) -> Tensor:
    """2-D tensor with unit inner stride (row pitch may exceed the width)."""
    if t.dim() != 2:
        raise RuntimeError("expected a 2-D tensor")
    if t.stride(1) != 1 and t.shape[1] > 1:
        t = t.contiguous()
    if t.shape[1] == 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _ld(t: Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(max(t.stride(0), t.shape[1]))


# --------------------------------------------------------------------------- #
# graph container
# --------------------------------------------------------------------------- #
class KnnGraph:
    """Fixed-width neighbour table: `nbr[N, W]` int32 (-1 padded), `deg[N]` int32.

    Row `i` lists the sources `j` of the edges `j -> i` in ascending distance;
    this is `edge_index` grouped by target, so it doubles as a CSR by target.
    """

    def __init__(self, nbr: Tensor, deg: Tensor, k: int):
        self.nbr, self.deg, self.k = nbr, deg, k
        self.n, self.width = int(nbr.shape[0]), int(nbr.shape[1])
        self._edge_index: Optional[Tensor] = None

    def to(self, device) -> "KnnGraph":
        return KnnGraph(self.nbr.to(device), self.deg.to(device), self.k)

    def edge_index(self) -> Tensor:
        """Materialise the PyG `[2, E]` int64 tensor (one host sync for E)."""
        if self._edge_index is None:
            _cuda(self.nbr)
            rowptr = torch.zeros(self.n + 1, dtype=torch.int64, device=self.nbr.device)
            torch.cumsum(self.deg, 0, out=rowptr[1:])
            e = int(rowptr[-1].item())
            ei = torch.empty(2, e, dtype=torch.int64, device=self.nbr.device)
            _call("gnb_table_to_edge_index", _ptr(self.nbr), _ptr(self.deg), _ptr(rowptr), self.n, self.width,
                  e, _ptr(ei), _stream())
            self._edge_index = ei
        return self._edge_index

    @staticmethod
    def from_edge_index(edge_index: Tensor, n: int, k: int = 0) -> "KnnGraph":
        """Foreign `edge_index` -> table (not on the hot path; syncs once for the max degree)."""
        _cuda(edge_index)
        src, dst = edge_index[0], edge_index[1]
        if src.numel() and not bool((dst[1:] >= dst[:-1]).all()):
            order = torch.argsort(dst, stable=True)
            src, dst = src[order], dst[order]
        rowptr = torch.searchsorted(dst.contiguous(), torch.arange(n + 1, device=dst.device))
        deg = (rowptr[1:] - rowptr[:-1])
        width = max(int(deg.max().item()) if n else 1, 1)
        nbr = torch.full((n, width), -1, dtype=torch.int32, device=dst.device)
        slot = torch.arange(src.numel(), device=dst.device) - rowptr[dst]
        nbr[dst, slot] = src.to(torch.int32)
        g = KnnGraph(nbr, deg.to(torch.int32), k or width)
        g._edge_index = torch.stack([src, dst])
        return g


# --------------------------------------------------------------------------- #
# graph building
# --------------------------------------------------------------------------- #
def batch_to_ptr(batch: Tensor, num_graphs: int) -> Tensor:
    _cuda(batch)
    batch = batch.contiguous()
    ptr = torch.empty(num_graphs + 1, dtype=torch.int64, device=batch.device)
    _call("gnb_batch_to_ptr", _ptr(batch), batch.numel(), num_graphs, _ptr(ptr), _stream())
    return ptr


_COLS_CACHE = {}


def _cols_tensor(cols: Sequence[int], device) -> Tensor:
    key = (tuple(cols), str(device))
    if key not in _COLS_CACHE:
        _COLS_CACHE[key] = torch.tensor(list(cols), dtype=torch.int32, device=device)
    return _COLS_CACHE[key]


def resolve_columns(features_subset, width: int) -> List[int]:
    if isinstance(features_subset, slice):
        return list(range(width))[features_subset]
    return [int(c) for c in features_subset]


def knn_table(x: Tensor, cols: Sequence[int], ptr: Tensor, k: int) -> KnnGraph:
    """Batched per-event kNN on columns `cols` of `x` (replaces `knn_graph`, layers.py:63-67)."""
    _cuda(x, ptr)
    x = _rowmajor(x.detach())
    if x.dtype != torch.float32:
        x = x.float()
    n = x.shape[0]
    nbr = torch.empty(n, k + 1, dtype=torch.int32, device=x.device)
    deg = torch.empty(n, dtype=torch.int32, device=x.device)
    _call("gnb_knn_table", _ptr(x), _ld(x), _ptr(_cols_tensor(cols, x.device)), len(cols), _ptr(ptr),
          ptr.numel() - 1, n, k, _ptr(nbr), _ptr(deg), _stream())
    return KnnGraph(nbr, deg, k)


def global_variables(x: Tensor, graph: KnnGraph, ptr: Tensor, n_pulses: Tensor,
                     x0_width: Optional[int] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """[mean(x) | h_x h_y h_z h_t | log10 n_pulses] per event (dynedge.py:266-293) and, optionally,
    x0 = [x | g[batch] | 0-pad] per node (dynedge.py:308-319)."""
    _cuda(x, ptr, n_pulses)
    x = _rowmajor(x.detach().float())
    nseg = ptr.numel() - 1
    nf = x.shape[1]
    g = torch.empty(nseg, nf + 5, dtype=torch.float32, device=x.device)
    x0 = None
    if x0_width is not None:
        x0 = torch.empty(x.shape[0], x0_width, dtype=torch.float32, device=x.device)
    npf = n_pulses.to(torch.float32).contiguous()
    _call("gnb_global_vars", _ptr(x), _ld(x), nf, _ptr(graph.nbr), _ptr(graph.deg), graph.width, _ptr(ptr), nseg,
          _ptr(npf), _ptr(g), _ptr(x0), 0 if x0 is None else x0_width, _stream())
    return g, x0


# --------------------------------------------------------------------------- #
# dense layers (fp32 SIMT backend; tensor-core backends hook in through `linear_*`)
# --------------------------------------------------------------------------- #
def _gemm_fwd(x: Tensor, w: Tensor, b: Optional[Tensor], y: Tensor, k: int, act: int, accumulate: bool) -> None:
    _call("gnb_linear_fwd_f32", _ptr(x), _ld(x), _ptr(w), _ld(w), _ptr(b), _ptr(y), _ld(y), x.shape[0], y.shape[1],
          k, act, 1 if accumulate else 0, _stream())


def _gemm_bwd_data(dz: Tensor, w: Tensor, dx: Tensor, k: int, accumulate: bool = False) -> None:
    _call("gnb_linear_bwd_data_f32", _ptr(dz), _ld(dz), _ptr(w), _ld(w), _ptr(dx), _ld(dx), dz.shape[0], dz.shape[1],
          k, 1 if accumulate else 0, _stream())


def _gemm_bwd_weight(dz: Tensor, x: Tensor, dw: Tensor, k: int) -> None:
    _call("gnb_linear_bwd_weight_f32", _ptr(dz), _ld(dz), _ptr(x), _ld(x), _ptr(dw), _ld(dw), dz.shape[0],
          dz.shape[1], k, _stream())


WGRAD_TC = os.environ.get("GNB_WGRAD_TC", "1") == "1"      # tcgen05 weight gradients in tf32 mode
WGRAD_SWAP = int(os.environ.get("GNB_WGRAD_SWAP", "0"))     # bring-up aid: swap LBO/SBO in the MN-major descriptors


def _gemm_bwd_weight_tc(dz: Tensor, x: Tensor, dw: Tensor, k: int) -> None:
    _call("gnb_linear_bwd_weight_tf32", _ptr(dz), _ld(dz), _ptr(x), _ld(x), _ptr(dw), _ld(dw), dz.shape[0],
          dz.shape[1], k, WGRAD_SWAP, _stream())


def _bias_act_backward(g: Tensor, y: Optional[Tensor], act: int, want_db: bool,
                       graph: Optional["KnnGraph"] = None, aggr: int = 0) -> Tuple[Tensor, Optional[Tensor]]:
    """dz = grow * act'(y) and db = colsum(dz) in ONE pass (gnb_act_bwd_colsum). With `graph`, `g` is the
    gradient of the aggregated node tensor and is broadcast to the node's valid edge slots first."""
    g = _rowmajor(g)
    cols = g.shape[1]
    rows = g.shape[0] if graph is None else graph.n * graph.width
    fast = cols % 4 == 0 and _ld(g) % 4 == 0 and g.data_ptr() % 16 == 0 and (y is None or (_ld(y) % 4 == 0 and
                                                                                     y.data_ptr() % 16 == 0))
    if not fast:   # odd widths only occur on tiny read-out tensors
        assert graph is None
        dz = g if act == ACT_NONE else g * (y > 0)
        return dz, (dz.sum(0) if want_db else None)
    if act == ACT_NONE and graph is None and not _tf32():
        return g, (_colsum(g) if want_db else None)
    dz = torch.empty(rows, cols, dtype=torch.float32, device=g.device)
    db = torch.zeros(cols, dtype=torch.float32, device=g.device) if want_db else None
    flags = act | (FLAG_ROUND_TF32 if _tf32() else 0)
    _call("gnb_act_bwd_colsum", _ptr(g), _ld(g), _ptr(y), 0 if y is None else _ld(y), rows, cols, _ptr(dz), _ld(dz),
          _ptr(db), flags, _ptr(None if graph is None else graph.deg), 1 if graph is None else graph.width, aggr,
          _stream())
    if _tf32():
        _mark_rounded(dz)
    return dz, db


def _round_pad(src: Tensor, dst_cols: Optional[int] = None) -> Tensor:
    """Fresh [rows, dst_cols] copy of `src` rounded to tf32 (cvt.rna), zero padded on the right."""
    src = _rowmajor(src)
    cols = src.shape[1]
    dst_cols = ((cols + 3) // 4) * 4 if dst_cols is None else dst_cols
    dst = torch.empty(src.shape[0], dst_cols, dtype=torch.float32, device=src.device)
    _call("gnb_round_pad_tf32", _ptr(src), _ld(src), src.shape[0], cols, _ptr(dst), dst_cols, dst_cols, _stream())
    return _mark_rounded(dst if dst_cols == cols else dst[:, :cols])    # logical width kept, pitch padded


def _tma_ok(t: Tensor) -> bool:
    return t.data_ptr() % 16 == 0 and _ld(t) % 4 == 0 and t.stride(1) == 1


def _tc_operand(t: Tensor) -> Tensor:
    """tf32-exact, TMA-addressable version of an activation matrix."""
    if _is_rounded(t) and _tma_ok(t):
        return t
    return _round_pad(t)


def _tma_operand(t: Tensor) -> Tensor:
    """TMA-addressable plain-fp32 version of an activation matrix (tf32x3: the GEMM splits it, or truncates it)."""
    if _tma_ok(t):
        return t
    cols = t.shape[1]
    padded = torch.zeros(t.shape[0], (cols + 3) // 4 * 4, dtype=torch.float32, device=t.device)
    padded[:, :cols] = t
    return padded[:, :cols]


def _tc_pack_weight_split(w: Tensor, offsets: Sequence[int], ks: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """(hi, lo) in the packed layout of `_tc_pack_weight`: hi = rna_tf32(w), lo = rna_tf32(w - hi)."""
    kbs = [(k + 31) // 32 * 32 for k in ks]
    hi = torch.empty(w.shape[0], sum(kbs), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    ldp, koff = hi.shape[1], 0
    for off, k, kb in zip(offsets, ks, kbs):
        src = w[:, off:off + k]
        _call("gnb_split_pad_tf32", _ptr(src), _ld(w), w.shape[0], k, ctypes.c_void_p(hi.data_ptr() + 4 * koff),
              ctypes.c_void_p(lo.data_ptr() + 4 * koff), ldp, kb, _stream())
        koff += kb
    return hi, lo


def _tc_linear_x3(parts: Sequence[Tensor], w_hi: Tensor, w_lo: Tensor, bias: Optional[Tensor], n_out: int, act: int) -> Tensor:
    rows, nparts = parts[0].shape[0], len(parts)
    y = torch.empty(rows, n_out, dtype=torch.float32, device=w_hi.device)
    xs = (ctypes.c_void_p * nparts)(*[p.data_ptr() for p in parts])
    lds = (ctypes.c_int64 * nparts)(*[_ld(p) for p in parts])
    ks = (ctypes.c_int32 * nparts)(*[p.shape[1] for p in parts])
    _call("gnb_linear_fwd_tf32x3", xs, lds, ks, nparts, _ptr(w_hi), _ptr(w_lo), w_hi.shape[1], _ptr(bias), _ptr(y), n_out,
          rows, n_out, act, _stream())
    return y


def _tc_pack_weight(w: Tensor, offsets: Sequence[int], ks: Sequence[int]) -> Tensor:
    """[n_out, sum ceil(k_p/32)*32]: part p's columns, rounded to tf32, at the 32-aligned running offset."""
    kbs = [(k + 31) // 32 * 32 for k in ks]
    packed = torch.empty(w.shape[0], sum(kbs), dtype=torch.float32, device=w.device)
    ldp, koff = packed.shape[1], 0
    for off, k, kb in zip(offsets, ks, kbs):
        src = w[:, off:off + k]
        _call("gnb_round_pad_tf32", _ptr(src), _ld(w), w.shape[0], k, ctypes.c_void_p(packed.data_ptr() + 4 * koff),
              ldp, kb, _stream())
        koff += kb
    return packed


def _tc_linear(parts: Sequence[Tensor], packed_w: Tensor, bias: Optional[Tensor], n_out: int, act: int,
               round_out: bool) -> Tensor:
    rows, nparts = parts[0].shape[0], len(parts)
    y = torch.empty(rows, n_out, dtype=torch.float32, device=packed_w.device)
    xs = (ctypes.c_void_p * nparts)(*[p.data_ptr() for p in parts])
    lds = (ctypes.c_int64 * nparts)(*[_ld(p) for p in parts])
    ks = (ctypes.c_int32 * nparts)(*[p.shape[1] for p in parts])
    _call("gnb_linear_fwd_tf32", xs, lds, ks, nparts, _ptr(packed_w), packed_w.shape[1], _ptr(bias), _ptr(y), n_out,
          rows, n_out, act, 1 if round_out else 0, _stream())
    return _mark_rounded(y) if round_out else y


def _colsum(a: Tensor) -> Tensor:
    out = torch.zeros(a.shape[1], dtype=torch.float32, device=a.device)
    _call("gnb_colsum", _ptr(a), _ld(a), a.shape[0], a.shape[1], _ptr(out), _stream())
    return out


def _dense_forward(parts: Sequence[Tensor], w: Tensor, b: Optional[Tensor], offsets: Sequence[int], act: int,
                   round_out: bool = True):
    """y = act(sum_p parts[p] @ w[:, off_p : off_p + K_p]^T + b). Returns (y, parts as consumed, used_tc).
    round_out: round y to tf32 (only needed when y itself feeds a tensor-core GEMM)."""
    rows, n_out = parts[0].shape[0], w.shape[0]
    tc = _tf32() and len(parts) <= 6 and rows > 0
    if tc and _split() and n_out <= 1024:
        parts = tuple(_tma_operand(p) for p in parts)
        w_hi, w_lo = _tc_pack_weight_split(w, offsets, [p.shape[1] for p in parts])
        y = _tc_linear_x3(parts, w_hi, w_lo, b, n_out, act)
    elif tc:
        parts = tuple(_tc_operand(p) for p in parts)
        packed = _tc_pack_weight(w, offsets, [p.shape[1] for p in parts])
        y = _tc_linear(parts, packed, b, n_out, act, round_out=round_out)
    else:
        y = torch.empty(rows, n_out, dtype=torch.float32, device=w.device)
        last = len(parts) - 1
        for i, (p, off) in enumerate(zip(parts, offsets)):
            _gemm_fwd(p, w[:, off:], b if i == last else None, y, p.shape[1],
                      act if i == last else ACT_NONE, accumulate=i > 0)
    return y, tuple(parts), tc


def _dense_backward(dz: Tensor, w: Tensor, parts: Sequence[Tensor], offsets: Sequence[int], need_dw: bool,
                    need_dparts: Sequence[bool], tc: bool):
    """dW = dz^T [parts] (into the packed column layout of w) and dparts[p] = dz @ w[:, off_p : off_p + K_p]."""
    dw = torch.zeros_like(w) if need_dw else None
    dzr = _tc_operand(dz) if tc else dz
    dparts: List[Optional[Tensor]] = []
    for p, off, need in zip(parts, offsets, need_dparts):
        kp = p.shape[1]
        if dw is not None:
            if tc and WGRAD_TC:       # tf32x3: the saved activation is plain fp32, the tensor core truncates it
                _gemm_bwd_weight_tc(dzr, _tma_operand(p) if _split() else _tc_operand(p), dw[:, off:], kp)
            else:
                _gemm_bwd_weight(dz, p, dw[:, off:], kp)
        if not need:
            dparts.append(None)
        elif tc:      # dx = dz W on the tensor cores: same kernel, W^T as the weight operand
            wt = w[:, off:off + kp].t().contiguous()                      # [kp, n_out]
            dparts.append(_tc_linear((dzr,), _tc_pack_weight(wt, (0,), (wt.shape[1],)), None, kp, ACT_NONE,
                                     round_out=False))
        else:
            dx = torch.empty(p.shape, dtype=torch.float32, device=p.device)
            _gemm_bwd_data(dz, w[:, off:], dx, kp)
            dparts.append(dx)
    return dw, dparts


class _MultiLinearAct(torch.autograd.Function):
    """y = act(sum_p parts[p] @ W[:, off_p : off_p + K_p]^T + b): a K-split Linear over several
    inputs, so the skip-concatenation of dynedge.py:328 is never materialised."""

    @staticmethod
    def forward(ctx, w: Tensor, b: Optional[Tensor], act: int, offsets: Tuple[int, ...], round_out: bool, *parts: Tensor):
        _cuda(w, b, *parts)
        w = _rowmajor(w)
        y, parts, ctx.tc = _dense_forward(tuple(_rowmajor(p) for p in parts), w, b, offsets, act, round_out)
        ctx.act, ctx.offsets = act, offsets
        ctx.has_bias = b is not None
        ctx.save_for_backward(w, y, *parts)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        w, y, *parts = ctx.saved_tensors
        dz, db = _bias_act_backward(gy.contiguous(), y, ctx.act, ctx.has_bias and ctx.needs_input_grad[1])
        dw, dparts = _dense_backward(dz, w, parts, ctx.offsets, ctx.needs_input_grad[0],
                                     [ctx.needs_input_grad[5 + i] for i in range(len(parts))], ctx.tc)
        return (dw, db, None, None, None, *dparts)


def linear_act(x: Tensor, w: Tensor, b: Optional[Tensor], act: int = ACT_NONE, round_out: bool = True) -> Tensor:
    """act(x @ w^T + b) (torch.nn.Linear + activation of dynedge.py:200-247)."""
    y = _MultiLinearAct.apply(w, b, act, (0,), round_out, x)
    return _mark_rounded(y) if (_fround() and round_out) else y


def multi_linear_act(parts: Sequence[Tensor], w: Tensor, b: Optional[Tensor], offsets: Sequence[int],
                     act: int = ACT_NONE) -> Tensor:
    y = _MultiLinearAct.apply(w, b, act, tuple(int(o) for o in offsets), True, *parts)
    return _mark_rounded(y) if _fround() else y


# --------------------------------------------------------------------------- #
# EdgeConv pieces
# --------------------------------------------------------------------------- #
class _EdgeHidden(torch.autograd.Function):
    """h[(i,s)] = act(P[i] + Q[nbr[i,s]]), the first Linear of the EdgeConv MLP hoisted to nodes:
    W1 [x_i ; x_j - x_i] + b1 = (W1a - W1b) x_i + b1 + W1b x_j."""

    @staticmethod
    def forward(ctx, pq: Tensor, graph: KnnGraph, act: int):
        _cuda(pq)
        pq = _rowmajor(pq)
        hdim = pq.shape[1] // 2
        h = torch.empty(graph.n * graph.width, hdim, dtype=torch.float32, device=pq.device)
        _call("gnb_edge_hidden_fwd", _ptr(pq), _ld(pq), hdim, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
              act | (FLAG_ROUND_TF32 if _fround() else 0), _ptr(h), _ld(h), _stream())
        ctx.graph, ctx.act = graph, act
        ctx.save_for_backward(h)
        return _mark_rounded(h) if _fround() else h

    @staticmethod
    def backward(ctx, gh: Tensor):
        (h,) = ctx.saved_tensors
        graph = ctx.graph
        gh = _rowmajor(gh.contiguous())
        hdim = h.shape[1]
        dpq = torch.zeros(graph.n, 2 * hdim, dtype=torch.float32, device=h.device)
        _call("gnb_edge_hidden_bwd", _ptr(gh), _ld(gh), _ptr(h), _ld(h), hdim, _ptr(graph.nbr), _ptr(graph.deg),
              graph.width, graph.n, ctx.act, _ptr(dpq), _ld(dpq), _stream())
        return dpq, None, None


def edge_hidden(pq: Tensor, graph: KnnGraph, act: int = ACT_RELU) -> Tensor:
    h = _EdgeHidden.apply(pq, graph, act)
    return _mark_rounded(h) if _fround() else h


class _EdgeCat(torch.autograd.Function):
    """u[(i,s)] = [x_i | x_j - x_i] (PyG EdgeConv.message input), padded-edge-list rows."""

    @staticmethod
    def forward(ctx, x: Tensor, graph: KnnGraph):
        _cuda(x)
        x = _rowmajor(x)
        c = x.shape[1]
        u = torch.empty(graph.n * graph.width, 2 * c, dtype=torch.float32, device=x.device)
        _call("gnb_edge_cat_fwd", _ptr(x), _ld(x), c, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
              _ptr(u), _ld(u), _stream())
        ctx.graph, ctx.c = graph, c
        return u

    @staticmethod
    def backward(ctx, du: Tensor):
        graph = ctx.graph
        du = _rowmajor(du.contiguous())
        dx = torch.zeros(graph.n, ctx.c, dtype=torch.float32, device=du.device)
        _call("gnb_edge_cat_bwd", _ptr(du), _ld(du), ctx.c, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
              _ptr(dx), _ld(dx), _stream())
        return dx, None


def edge_cat(x: Tensor, graph: KnnGraph) -> Tensor:
    return _EdgeCat.apply(x, graph)


class _EdgeAggregate(torch.autograd.Function):
    """y[i] = AGG_{s < deg[i]} m[(i,s)] with AGG in add / mean / max (arg-routed backward)."""

    @staticmethod
    def forward(ctx, m: Tensor, graph: KnnGraph, aggr: int):
        _cuda(m)
        m = _rowmajor(m)
        c = m.shape[1]
        y = torch.empty(graph.n, c, dtype=torch.float32, device=m.device)
        arg = torch.empty(graph.n, c, dtype=torch.int8, device=m.device) if aggr == 2 else None
        _call("gnb_edge_aggregate_fwd", _ptr(m), _ld(m), c, _ptr(graph.deg), graph.width, graph.n,
              aggr | (FLAG_ROUND_TF32 if _fround() else 0), _ptr(y), _ld(y), _ptr(arg), _stream())
        ctx.graph, ctx.aggr, ctx.c = graph, aggr, c
        ctx.save_for_backward(arg) if arg is not None else None
        return _mark_rounded(y) if _fround() else y

    @staticmethod
    def backward(ctx, gy: Tensor):
        graph = ctx.graph
        arg = ctx.saved_tensors[0] if ctx.aggr == 2 else None
        gy = _rowmajor(gy.contiguous())
        gm = torch.empty(graph.n * graph.width, ctx.c, dtype=torch.float32, device=gy.device)
        _call("gnb_edge_aggregate_bwd", _ptr(gy), _ld(gy), ctx.c, _ptr(graph.deg), graph.width, graph.n, ctx.aggr,
              _ptr(arg), _ptr(gm), _ld(gm), _stream())
        return gm, None, None


def edge_aggregate(m: Tensor, graph: KnnGraph, aggr: str = "add") -> Tensor:
    y = _EdgeAggregate.apply(m, graph, AGGR[aggr])
    return _mark_rounded(y) if _fround() else y


class _EdgeConvHoisted(torch.autograd.Function):
    """y_i = AGG_s act(W2 act(P_i + Q_{nbr[i,s]}) + b2): the whole per-edge part of a DynEdge EdgeConv as one
    autograd node, so its backward runs as fused kernels: (aggregate-bwd + ReLU-bwd + bias-grad) in one pass,
    dW2 and dh on the tensor cores, (ReLU-bwd + dP segment-sum + dQ scatter) in one pass."""

    @staticmethod
    def forward(ctx, pq: Tensor, w2: Tensor, b2: Optional[Tensor], graph: KnnGraph, aggr: int):
        _cuda(pq, w2, b2)
        pq, w2 = _rowmajor(pq), _rowmajor(w2)
        hdim = pq.shape[1] // 2
        rnd = FLAG_ROUND_TF32 if _fround() else 0
        h = torch.empty(graph.n * graph.width, hdim, dtype=torch.float32, device=pq.device)
        _call("gnb_edge_hidden_fwd", _ptr(pq), _ld(pq), hdim, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
              ACT_RELU | rnd, _ptr(h), _ld(h), _stream())
        if _fround():
            _mark_rounded(h)
        m, (h,), ctx.tc = _dense_forward((h,), w2, b2, (0,), ACT_RELU, round_out=False)    # summed in fp32
        c = m.shape[1]
        y = torch.empty(graph.n, c, dtype=torch.float32, device=pq.device)
        _call("gnb_edge_aggregate_fwd", _ptr(m), _ld(m), c, _ptr(graph.deg), graph.width, graph.n, aggr | rnd, _ptr(y),
              _ld(y), _ptr(None), _stream())
        ctx.graph, ctx.aggr, ctx.has_bias = graph, aggr, b2 is not None
        ctx.save_for_backward(h, m, w2)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        h, m, w2 = ctx.saved_tensors
        graph = ctx.graph
        dz, db2 = _bias_act_backward(gy.contiguous(), m, ACT_RELU, ctx.has_bias and ctx.needs_input_grad[2], graph,
                                     ctx.aggr)
        dw2, (dh,) = _dense_backward(dz, w2, (h,), (0,), ctx.needs_input_grad[1], (ctx.needs_input_grad[0],), ctx.tc)
        dpq = None
        if dh is not None:
            hdim = h.shape[1]
            dpq = torch.zeros(graph.n, 2 * hdim, dtype=torch.float32, device=h.device)
            _call("gnb_edge_hidden_bwd", _ptr(dh), _ld(dh), _ptr(h), _ld(h), hdim, _ptr(graph.nbr), _ptr(graph.deg),
                  graph.width, graph.n, ACT_RELU, _ptr(dpq), _ld(dpq), _stream())
        return dpq, dw2, db2, None, None


class _EdgeConvHoistedMax(torch.autograd.Function):
    """y_i = max_s act2(W2 act1(P_i + Q_{nbr[i,s]}) + b2): the per-edge part of an EdgeConv with max aggregation (EdgeConvTito,
    reference layers.py:72-114 + dynedge_kaggle_tito.py:157-162) on the tensor cores. Forward: hidden-layer kernel ->
    tcgen05 GEMM whose epilogue applies act2, takes the maximum over the k slots of every node and records the winning slot
    (the [E, C] message tensor is never stored). Backward: the arg-routed kernel rebuilds dz [E, C] (one non-zero per node and
    channel), dW2 and dh run on the tensor cores, the hidden-layer backward scatters dh to dPQ."""

    @staticmethod
    def forward(ctx, pq: Tensor, w2: Tensor, b2: Optional[Tensor], graph: KnnGraph, act1: int, act2: int):
        _cuda(pq, w2, b2)
        pq, w2 = _rowmajor(pq), _rowmajor(w2)
        hdim, c_out = pq.shape[1] // 2, w2.shape[0]
        rnd = FLAG_ROUND_TF32 if _fround() else 0
        h = torch.empty(graph.n * graph.width, hdim, dtype=torch.float32, device=pq.device)
        _call("gnb_edge_hidden_fwd", _ptr(pq), _ld(pq), hdim, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
              act1 | rnd, _ptr(h), _ld(h), _stream())
        y = torch.empty(graph.n, c_out, dtype=torch.float32, device=pq.device)
        arg = torch.empty(graph.n, c_out, dtype=torch.int8, device=pq.device)
        bias = None if b2 is None else b2.detach()
        if _split():
            w_hi, w_lo = _tc_pack_weight_split(w2.detach(), (0,), (hdim,))
            _call("gnb_edge_linear_aggmax_fwd_tf32x3", _ptr(h), _ld(h), hdim, _ptr(w_hi), _ptr(w_lo), w_hi.shape[1], _ptr(bias),
                  _ptr(graph.deg), graph.n, c_out, act2, _ptr(y), c_out, _ptr(arg), c_out, _stream())
        else:
            _mark_rounded(h)
            w2p = _tc_pack_weight(w2.detach(), (0,), (hdim,))
            _call("gnb_edge_linear_aggmax_fwd_tf32", _ptr(h), _ld(h), hdim, _ptr(w2p), w2p.shape[1], _ptr(bias),
                  _ptr(graph.deg), graph.n, c_out, act2, 1, _ptr(y), c_out, _ptr(arg), c_out, _stream())
            _mark_rounded(y)
        ctx.graph, ctx.act1, ctx.act2, ctx.has_bias = graph, act1, act2, b2 is not None
        ctx.save_for_backward(h, arg, w2)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        h, arg, w2 = ctx.saved_tensors
        graph = ctx.graph
        gy = _rowmajor(gy.contiguous())
        c_out = w2.shape[0]
        dz = torch.empty(graph.n * graph.width, c_out, dtype=torch.float32, device=h.device)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        db2 = torch.zeros(c_out, dtype=torch.float32, device=h.device) if want_db else None
        _call("gnb_edge_argmax_bwd", _ptr(gy), _ld(gy), _ptr(arg), c_out, c_out, graph.width, graph.n, ctx.act2,
              FLAG_ROUND_TF32, _ptr(dz), c_out, _ptr(db2), _stream())
        _mark_rounded(dz)
        dw2, (dh,) = _dense_backward(dz, w2, (h,), (0,), ctx.needs_input_grad[1], (ctx.needs_input_grad[0],), True)
        dpq = None
        if dh is not None:
            hdim = h.shape[1]
            dpq = torch.zeros(graph.n, 2 * hdim, dtype=torch.float32, device=h.device)
            _call("gnb_edge_hidden_bwd", _ptr(dh), _ld(dh), _ptr(h), _ld(h), hdim, _ptr(graph.nbr), _ptr(graph.deg),
                  graph.width, graph.n, ctx.act1, _ptr(dpq), _ld(dpq), _stream())
        return dpq, dw2, db2, None, None, None


MAX_EPILOGUE = True       # False: EdgeConv layers with aggr = "max" keep the unfused route (timing comparisons)


def edgeconv_hoisted_max_ok(graph: KnnGraph, hdim: int, c_out: int) -> bool:
    """The tensor-core max-aggregation route exists for k = 8 tables (width 9) in the tensor-core precision modes."""
    return MAX_EPILOGUE and _tf32() and graph.width == 9 and hdim % 4 == 0 and 4 <= hdim <= 2048 and 1 <= c_out <= 512 and graph.n > 0


def edgeconv_hoisted_max(pq: Tensor, w2: Tensor, b2: Optional[Tensor], graph: KnnGraph, act1: int, act2: int) -> Tensor:
    """Per-edge half of an EdgeConv `Linear, act1, Linear, act2` with aggr = "max" whose first Linear was hoisted to nodes
    (pq = [P | Q]); act in ACT_NONE / ACT_RELU / ACT_LEAKY."""
    y = _EdgeConvHoistedMax.apply(pq, w2, b2, graph, act1, act2)
    return _mark_rounded(y) if _fround() else y


FUSED_EDGECONV = os.environ.get("GNB_FUSED_EDGECONV", "1") == "1"
# inference route of the executor for k = 8 graphs: "split" = hidden-layer kernel + aggregating CTA-pair GEMM (measured
# faster on B200), "fused" = the single fused EdgeConv kernel (gather + MLP + aggregation, nothing per-edge in HBM)
INFERENCE_ROUTE = os.environ.get("GNB_INFERENCE_ROUTE", "split")


def set_edgeconv_variant(v: int) -> None:
    """0 auto, 1 single-CTA fused EdgeConv kernel, 2 CTA-pair (cta_group::2) kernel."""
    _lib.check(_lib.load().gnb_edgeconv_set_variant(v), "gnb_edgeconv_set_variant")


def edgeconv_fused_forward(pq: Tensor, w2: Tensor, b2: Optional[Tensor], graph: KnnGraph, aggr: str = "add") -> Tensor:
    """Inference-only fused EdgeConv (gather + hidden ReLU + contraction + bias/ReLU + aggregation in ONE tcgen05
    kernel; no [E, H] / [E, C] tensors in HBM). tf32 mode."""
    _cuda(pq, w2, b2)
    pq = _rowmajor(pq.detach())
    w2 = _rowmajor(w2.detach())
    hdim, c_out = w2.shape[1], w2.shape[0]
    w2p = _tc_pack_weight(w2, (0,), (hdim,))
    y = torch.empty(graph.n, c_out, dtype=torch.float32, device=pq.device)
    _call("gnb_edgeconv_fused_fwd_tf32", _ptr(pq), _ld(pq), hdim, _ptr(graph.nbr), _ptr(graph.deg), graph.width, graph.n,
          _ptr(w2p), w2p.shape[1], _ptr(None if b2 is None else b2.detach()), c_out, AGGR[aggr], 1, _ptr(y), c_out, _stream())
    return _mark_rounded(y)


def edgeconv_hoisted(pq: Tensor, w2: Tensor, b2: Optional[Tensor], graph: KnnGraph, aggr: str = "add") -> Tensor:
    """Per-edge half of EdgeConv for an MLP `Linear, ReLU, Linear, ReLU` whose first Linear was hoisted to nodes
    (pq = [P | Q]); aggr in add / mean. (max keeps the unfused route: it needs the arg-routed backward.)"""
    needs_grad = torch.is_grad_enabled() and (pq.requires_grad or w2.requires_grad or (b2 is not None and b2.requires_grad))
    if _fround() and FUSED_EDGECONV and not needs_grad and w2.shape[1] <= 352 and w2.shape[1] % 4 == 0 and graph.width <= 32:
        return edgeconv_fused_forward(pq, w2, b2, graph, aggr)
    y = _EdgeConvHoisted.apply(pq, w2, b2, graph, AGGR[aggr])
    return _mark_rounded(y) if _fround() else y


# --------------------------------------------------------------------------- #
# global pooling (dynedge.py:251-264)
# --------------------------------------------------------------------------- #
class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, ptr: Tensor, schemes: Tuple[int, ...]):
        _cuda(x, ptr)
        x = _rowmajor(x)
        nseg, c, npool = ptr.numel() - 1, x.shape[1], len(schemes)
        out = torch.empty(nseg, npool * c, dtype=torch.float32, device=x.device)
        arg = torch.empty(nseg, npool * c, dtype=torch.int32, device=x.device)
        sch = (ctypes.c_int32 * npool)(*schemes)
        _call("gnb_segment_pool_fwd", _ptr(x), _ld(x), c, _ptr(ptr), nseg, sch, npool, _ptr(out), _ptr(arg), _stream())
        ctx.schemes, ctx.n, ctx.c = schemes, x.shape[0], c
        ctx.save_for_backward(arg, ptr)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, gout: Tensor, _garg):
        arg, ptr = ctx.saved_tensors
        gout = gout.contiguous()
        npool = len(ctx.schemes)
        gx = torch.empty(ctx.n, ctx.c, dtype=torch.float32, device=gout.device)
        sch = (ctypes.c_int32 * npool)(*ctx.schemes)
        _call("gnb_segment_pool_bwd", _ptr(gout), gout.shape[1], _ptr(arg), ctx.c, _ptr(ptr), ptr.numel() - 1, ctx.n, sch, npool,
              _ptr(gx), _ld(gx), _stream())
        return gx, None, None


def segment_pool(x: Tensor, ptr: Tensor, schemes: Sequence[str], return_arg: bool = False):
    """cat_p scatter_<p>(x, batch) -> [B, P*C] in the caller's scheme order."""
    out, arg = _SegmentPool.apply(x, ptr, tuple(POOL[s] for s in schemes))
    return (out, arg) if return_arg else out


# --------------------------------------------------------------------------- #
# native step executor (csrc/dynedge_exec.cu): DynEdge.forward / backward from one C call each
# --------------------------------------------------------------------------- #
USE_EXECUTOR = os.environ.get("GNB_EXECUTOR", "1") == "1"
# When True and every parameter already owns a contiguous fp32 `.grad` (e.g. the views of
# `distributed.FlatGradAllReduce`), the executor's backward accumulates straight into those buffers and hands
# autograd `None`, saving one zero-fill and one add kernel per parameter tensor. Off by default because autograd
# hooks on the parameters would not observe these gradients.
ACCUMULATE_INTO_GRAD = False


# Workspace pool: the executor needs one multi-GB scratch buffer per step. Allocating it through the caching
# allocator every call makes step time depend on allocator state (a 150 ms cudaMalloc hiccup was measured);
# buffers are therefore recycled here (grow-only, same-stream use assumed, at most 4 kept per device).
_WS_POOL: dict = {}


def _ws_acquire(nbytes: int, device) -> Tensor:
    pool = _WS_POOL.setdefault(str(device), [])
    best = -1
    for i, t in enumerate(pool):
        if t.numel() >= nbytes and (best < 0 or t.numel() < pool[best].numel()):
            best = i
    if best >= 0:
        return pool.pop(best)
    return torch.empty(int(nbytes * 1.2) + 256, dtype=torch.uint8, device=device)


def _ws_release(t: Tensor) -> None:
    pool = _WS_POOL.setdefault(str(t.device), [])
    pool.append(t)
    if len(pool) > 4:
        pool.sort(key=lambda b: b.numel())
        pool.pop(0)


def _copy_cfg(cfg):
    dup = type(cfg)()
    ctypes.memmove(ctypes.byref(dup), ctypes.byref(cfg), ctypes.sizeof(cfg))
    return dup


class _DynEdgeExec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, graph: KnnGraph, ptr: Tensor, n_pulses: Tensor, x: Tensor, cols_dev: Tensor, out_cols: int,
                out_per_event: bool, record: Optional[dict], training: int, *params: Tensor):
        _cuda(x, ptr, n_pulses, *params)
        lib = _lib.load()
        x = _rowmajor(x.detach().float())
        n, nseg = x.shape[0], ptr.numel() - 1
        cfg = _copy_cfg(cfg)      # the ctx keeps ITS OWN copy: a later set_precision() / flag change cannot alter the layout
        cfg.precision = 2 if _split() else (1 if _tf32() else 0)
        cfg.flags = (0 if FUSED_EDGECONV else 1) | (2 if INFERENCE_ROUTE == "split" else 0) | (4 if STORE_DZ else 0) | (8 if UNFUSED_FORWARD else 0) | (16 if SLOTS9 else 0)
        nbytes = -2
        if PRECISION in ("bf16", "bf16x3", "mixed16", "f16"):     # per-edge tensors as 16-bit planes where the configuration has the k = 8 route
            base = cfg.precision
            cfg.precision = {"bf16": 3, "bf16x3": 4, "mixed16": 5, "f16": 6}[PRECISION]
            nbytes = lib.gnb_dynedge_workspace_bytes(ctypes.byref(cfg), n, nseg, graph.width, training)
            if nbytes == -2:
                cfg.precision = base
        if nbytes == -2:
            nbytes = lib.gnb_dynedge_workspace_bytes(ctypes.byref(cfg), n, nseg, graph.width, training)
        if nbytes < 0:
            raise RuntimeError(f"gnb_dynedge_workspace_bytes: unsupported configuration [{nbytes}]")
        ws = _ws_acquire(int(nbytes), x.device)
        out = torch.empty(nseg if out_per_event else n, out_cols, dtype=torch.float32, device=x.device)
        npf = n_pulses.to(torch.float32).contiguous()
        plist = [p.detach().contiguous() for p in params]
        parr = (ctypes.c_void_p * len(plist))(*[p.data_ptr() for p in plist])
        _call("gnb_dynedge_forward", ctypes.byref(cfg), parr, _ptr(x), _ld(x), _ptr(ptr), _ptr(npf), _ptr(graph.nbr),
              _ptr(graph.deg), graph.width, _ptr(cols_dev), n, nseg, _ptr(ws), int(nbytes), _ptr(out),
              training, _stream())
        if record is not None:       # test hook: views of the per-layer outputs / recomputed graphs inside `ws`
            offs = (ctypes.c_int64 * (3 * cfg.n_conv))()
            _lib.check(lib.gnb_dynedge_layout(ctypes.byref(cfg), n, nseg, graph.width, training, offs), "layout")
            ys, graphs = [], [graph]
            for li in range(cfg.n_conv):
                c = cfg.conv_out[li]
                ys.append(ws[offs[3 * li]: offs[3 * li] + 4 * n * c].view(torch.float32).view(n, c))
                if offs[3 * li + 1] >= 0:
                    w = cfg.k + 1
                    nbr = ws[offs[3 * li + 1]: offs[3 * li + 1] + 4 * n * w].view(torch.int32).view(n, w)
                    deg = ws[offs[3 * li + 2]: offs[3 * li + 2] + 4 * n].view(torch.int32)
                    graphs.append(KnnGraph(nbr, deg, cfg.k))
            record["ys"], record["graphs"] = ys, graphs
            ng = cfg.nb_inputs + 5
            record["g"] = ws[: 4 * nseg * ng].view(torch.float32).view(nseg, ng)
            node_w = cfg.nb_inputs + (0 if cfg.globals_after_pooling else ng)
            x0_ld = (node_w + 31) // 32 * 32
            x0_off = (4 * nseg * ng + 255) // 256 * 256
            record["x0"] = ws[x0_off: x0_off + 4 * n * x0_ld].view(torch.float32).view(n, x0_ld)[:, :node_w]
        ctx.cfg, ctx.graph, ctx.nbytes, ctx.n, ctx.nseg, ctx.training = cfg, graph, int(nbytes), n, nseg, training
        if training:
            ctx.save_for_backward(ws, ptr, *params)
        elif record is None:
            _ws_release(ws)          # inference: nothing is kept, the buffer can serve the next call right away
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        if not ctx.training:
            raise RuntimeError("DynEdge executor: backward requested after an inference-mode forward")
        if getattr(ctx, "consumed", False):
            raise RuntimeError("DynEdge executor: the workspace of this forward was already consumed by a backward pass "
                               "(retain_graph / double backward are not supported on the executor route)")
        ctx.consumed = True
        ws, ptr, *params = ctx.saved_tensors
        gout = gout.contiguous().float()
        direct = ACCUMULATE_INTO_GRAD and all(
            p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.shape == p.shape
            for p in params)
        grads = [p.grad for p in params] if direct else \
            [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in params]
        garr = (ctypes.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
        graph = ctx.graph
        _call("gnb_dynedge_backward", ctypes.byref(ctx.cfg), garr, _ptr(ptr), _ptr(graph.nbr), _ptr(graph.deg), graph.width,
              ctx.n, ctx.nseg, _ptr(ws), ctx.nbytes, _ptr(gout), _stream())
        _ws_release(ws)
        if direct:
            return (None,) * (10 + len(params))
        return (None,) * 10 + tuple(grads)


def dynedge_execute(cfg, graph: KnnGraph, ptr: Tensor, n_pulses: Tensor, x: Tensor, cols: Sequence[int], out_cols: int,
                    out_per_event: bool, params: Sequence[Tensor], record: Optional[dict] = None) -> Tensor:
    training = 1 if (torch.is_grad_enabled() and any(p.requires_grad for p in params)) else 0
    return _DynEdgeExec.apply(cfg, graph, ptr, n_pulses, x, _cols_tensor(cols, x.device), out_cols, out_per_event,
                              record, training, *params)


# --------------------------------------------------------------------------- #
# fused task heads + losses (csrc/task_heads.cu): energy (log-cosh on log10) + direction (vMF-3D) in two kernels
# --------------------------------------------------------------------------- #
class _TaskHeadsLoss(torch.autograd.Function):
    """loss[2] = (mean log-cosh(log10 E_pred - log10 E), mean vMF-3D NLL); also returns the predictions (detached)."""

    @staticmethod
    def forward(ctx, feat: Tensor, we: Tensor, be: Tensor, wd: Tensor, bd: Tensor, energy: Tensor, direction: Tensor):
        _cuda(feat, we, be, wd, bd, energy, direction)
        feat = _rowmajor(feat.detach().float())
        we_, be_, wd_, bd_ = (t.detach().float().contiguous() for t in (we, be, wd, bd))
        energy = energy.detach().float().contiguous()
        direction = direction.detach().float().reshape(-1, 3).contiguous()
        nev, hdim = feat.shape
        pred_e = torch.empty(nev, device=feat.device)
        pred_d = torch.empty(nev, 4, device=feat.device)
        dz = torch.empty(nev, 4, device=feat.device)
        loss = torch.zeros(2, device=feat.device)
        _call("gnb_task_heads_fwd", _ptr(feat), _ld(feat), hdim, _ptr(we_), _ptr(be_), _ptr(wd_), _ptr(bd_), _ptr(energy),
              _ptr(direction), nev, _ptr(pred_e), _ptr(pred_d), _ptr(dz), _ptr(loss), _stream())
        ctx.save_for_backward(feat, we_, wd_, dz)
        ctx.params = (we, be, wd, bd)
        ctx.mark_non_differentiable(pred_e, pred_d)
        return loss, pred_e, pred_d

    @staticmethod
    def backward(ctx, gloss: Tensor, _ge, _gd):
        feat, we_, wd_, dz = ctx.saved_tensors
        nev, hdim = feat.shape
        we, be, wd, bd = ctx.params
        # d(sum of both losses): the two loss terms share one upstream scale only if their gradients are equal, which
        # is the case for loss.sum() / loss[0] + loss[1]; general weights are applied per head below
        g = gloss.contiguous().float()
        if not bool(ctx.needs_input_grad[0]) and not any(ctx.needs_input_grad[1:5]):
            return (None,) * 7
        dzs = dz * torch.stack([g[0], g[1], g[1], g[1]]).unsqueeze(0)      # per-head upstream gradient
        dfeat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
        direct = ACCUMULATE_INTO_GRAD and all(
            p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.shape == p.shape
            for p in (we, be, wd, bd))
        grads = [p.grad for p in (we, be, wd, bd)] if direct else \
            [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in (we, be, wd, bd)]
        _call("gnb_task_heads_bwd", _ptr(feat), _ld(feat), hdim, _ptr(we_), _ptr(wd_), _ptr(dzs), _ptr(None), nev,
              _ptr(dfeat), hdim, _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]), _ptr(grads[3]), _stream())
        if direct:
            return (dfeat, None, None, None, None, None, None)
        return (dfeat, grads[0], grads[1], grads[2], grads[3], None, None)


def task_heads_loss(feat: Tensor, we: Tensor, be: Tensor, wd: Tensor, bd: Tensor, energy: Tensor, direction: Tensor):
    """(loss[2], pred_energy[B], pred_direction[B, 4]) of the fused energy + direction heads."""
    return _TaskHeadsLoss.apply(feat, we, be, wd, bd, energy, direction)
