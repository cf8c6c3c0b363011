"""Time the tcgen05 Linear kernel alone on the EdgeConv shapes of the training step, with the profiling switches of
gnb_linear_set_debug (which role bounds the kernel?). Run on a GPU box: python scripts/linear_probe.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphnet_b200 import ops, _lib  # noqa: E402

ops.set_precision("tf32")
lib = _lib.load()
dev = torch.device("cuda", 0)
N = 79261
ROWS = N * 9


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(reps))
    return ts[len(ts) // 2]


def linear_case(rows, k, n_out, label):
    x = ops._mark_rounded(torch.randn(rows, k, device=dev).round_())
    w = torch.randn(n_out, k, device=dev).round_()
    b = torch.randn(n_out, device=dev)
    pw = ops._tc_pack_weight(w, [0], [k])
    out = []
    for dbg in (0, 1, 2, 3, 15):
        lib.gnb_linear_set_debug(dbg)
        t = timed(lambda: ops._tc_linear([x], pw, b, n_out, 1, True))
        out.append(f"dbg{dbg}={t:.0f}")
    lib.gnb_linear_set_debug(0)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    for dbg in (0, 1, 15):
        lib.gnb_linear_set_debug(dbg)
        lib.gnb_linear_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
        ops._tc_linear([x], pw, b, n_out, 1, True)
        torch.cuda.synchronize()
        lib.gnb_linear_set_profile_buffer(None)
        v = prof.tolist()
        print(f"   dbg{dbg} cycles: producer wait-empty {v[0]} total {v[2]} | mma wait-full {v[3]} wait-tmem-empty {v[4]} "
              f"total {v[5]} | epilogue wait-tmem-full {v[6]} tmem-ld {v[7]} total {v[8]}", flush=True)
    lib.gnb_linear_set_debug(0)
    gf = 2.0 * rows * k * n_out
    print(f"{label}: rows={rows} k={k} n_out={n_out} us: " + " ".join(out) +
          f" | {gf / 1e6 / float(out[0].split('=')[1]):.0f} TFLOP/s", flush=True)


def agg_case(n, k, n_out, label):
    rows = n * 9
    h = torch.randn(rows, k, device=dev).round_()
    w = torch.randn(n_out, k, device=dev).round_()
    b = torch.randn(n_out, device=dev)
    pw = ops._tc_pack_weight(w, [0], [k])
    deg = torch.full((n,), 8, dtype=torch.int32, device=dev)
    y = torch.empty(n, n_out, device=dev)
    ntile = (n + 13) // 14
    mask = torch.empty(ntile * n_out * 4, dtype=torch.int32, device=dev)
    out = []
    for dbg in (0, 2):
        lib.gnb_linear_set_debug(dbg)

        def run():
            ops._call("gnb_edge_linear_agg_fwd_tf32", ops._ptr(h), k, k, ops._ptr(pw), pw.shape[1], ops._ptr(b),
                      ops._ptr(deg), n, n_out, 1, ops._ptr(y), n_out, ops._ptr(mask), ops._stream())
        out.append(f"dbg{dbg}={timed(run):.0f}")
    lib.gnb_linear_set_debug(0)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.gnb_linear_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gnb_linear_set_profile_buffer(None)
    v = prof.tolist()
    print(f"   agg cycles: producer wait-empty {v[0]} total {v[2]} | mma wait-full {v[3]} wait-tmem-empty {v[4]} "
          f"total {v[5]} | epilogue wait-tmem-full {v[6]} arrive {v[7]} total {v[8]} deg-loads {v[9]} tmem-ld {v[10]}", flush=True)
    print(f"{label}: n={n} k={k} n_out={n_out} us: " + " ".join(out), flush=True)


def scatter_case(n, c_out, hdim, label, pitch=None):
    rows = n * 9
    pitch = pitch or c_out
    dz = torch.randn(rows, pitch, device=dev).round_()[:, :c_out]
    kpad = (c_out + 31) // 32 * 32
    wt = torch.randn(hdim, kpad, device=dev).round_()
    mld = 4 * ((hdim + 127) // 128)
    hmask = torch.randint(-2**31, 2**31 - 1, ((n + 13) // 14 * 126, mld), device=dev, dtype=torch.int32)
    # neighbours: random nodes within +-64 of the target (kNN-like locality), degree 8
    base = torch.arange(n, device=dev).unsqueeze(1)
    nbr = (base + torch.randint(-64, 65, (n, 9), device=dev)).clamp_(0, n - 1).int()
    nbr[:, 8] = -1
    dpq = torch.zeros(n, 2 * hdim, device=dev)
    out = []
    for dbg in (0, 2, 64, 66):
        lib.gnb_linear_set_debug(dbg)

        def run():
            ops._call("gnb_edge_hidden_dgrad_scatter_tf32", ops._ptr(dz), pitch, c_out, ops._ptr(wt), kpad, ops._ptr(hmask), mld,
                      hdim, ops._ptr(nbr), n, ops._ptr(dpq), 2 * hdim, ops._stream())
        out.append(f"dbg{dbg}={timed(run):.0f}")
    lib.gnb_linear_set_debug(0)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.gnb_linear_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gnb_linear_set_profile_buffer(None)
    v = prof.tolist()
    print(f"   scatter cycles (cluster 0): producer wait-empty {v[0]} total {v[2]} | mma wait-full {v[3]} wait-tmem-empty {v[4]} "
          f"total {v[5]} | epilogue wait-tmem-full {v[6]} arrive {v[7]} total {v[8]}", flush=True)
    print(f"{label}: n={n} c_out={c_out} hdim={hdim} us: " + " ".join(out), flush=True)


import sys as _s
for variant, resident in (((1, 0), (2, 0)) if not (len(_s.argv) > 1 and _s.argv[1] in ("resident", "pitch", "groups", "dual")) else ()):
    lib.gnb_linear_set_variant(variant)
    lib.gnb_linear_set_pair_resident(resident)
    print("== variant", variant, "(1 single-CTA, 2 CTA pair), resident weights from", resident, "stages", flush=True)
    scatter_case(N, 256, 336, "dgrad + scatter epilogue")
    scatter_case(N, 256, 128, "layer-1 dgrad + scatter")
    agg_case(N, 336, 256, "edge GEMM2 fwd (aggregating)")
    linear_case(ROWS, 336, 256, "edge GEMM2 fwd (plain)")
lib.gnb_linear_set_variant(0)
lib.gnb_linear_set_pair_resident(0)
if len(_s.argv) > 1 and _s.argv[1] == "scatter":
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "dual":
    for variant in (2, 3):
        lib.gnb_linear_set_variant(variant)
        print("== variant", variant, "(2 = one cluster set per channel group, 3 = dual-group kernel)", flush=True)
        scatter_case(N, 256, 336, "dgrad + scatter epilogue")
    lib.gnb_linear_set_variant(0)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "groups":
    # how much of the 336-channel launch is the second (80-channel) group's re-stream of dz?
    lib.gnb_linear_set_variant(2)
    for hd in (256, 336, 128, 80):
        scatter_case(N, 256, hd, f"dgrad + scatter, hdim {hd}")
    lib.gnb_linear_set_variant(0)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "pitch":
    lib.gnb_linear_set_variant(2)
    for pitch in (256, 272, 288, 320, 336):
        scatter_case(N, 256, 336, f"dgrad + scatter, dz pitch {pitch}", pitch=pitch)
    for pitch in (256, 272):
        scatter_case(N, 256, 128, f"layer-1 dgrad + scatter, dz pitch {pitch}", pitch=pitch)
    lib.gnb_linear_set_variant(0)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "resident":
    for resident in (0, 3, 2):
        lib.gnb_linear_set_variant(2)
        lib.gnb_linear_set_pair_resident(resident)
        print("== CTA pair, resident weights from", resident, "stages", flush=True)
        scatter_case(N, 256, 336, "dgrad + scatter epilogue")
        scatter_case(N, 256, 128, "layer-1 dgrad + scatter")
        agg_case(N, 336, 256, "edge GEMM2 fwd (aggregating)")
        linear_case(N, 256, 672, "PQ linear")
    lib.gnb_linear_set_variant(0)
    lib.gnb_linear_set_pair_resident(0)
    _s.exit(0)
linear_case(ROWS, 336, 256, "edge GEMM2 fwd (plain)")
linear_case(ROWS, 256, 336, "edge GEMM2 dgrad")
linear_case(ROWS, 256, 128, "layer-1 dgrad")
linear_case(N, 256, 672, "PQ linear")
agg_case(N, 336, 256, "edge GEMM2 fwd (aggregating)")
