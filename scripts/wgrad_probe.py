"""wgrad time per row: full edge list (operands from HBM) vs a chunk small enough to stay in the 126 MB L2."""
import sys
import torch
sys.path.insert(0, ".")
from graphnet_b200 import ops, _lib
ops.set_precision("tf32")
lib = _lib.load()
dev = torch.device("cuda", 0)
ROWS = 713349


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return b.elapsed_time(e) / reps * 1e3


dz = torch.randn(ROWS, 256, device=dev).round_()
h = torch.randn(ROWS, 336, device=dev).round_()
dw = torch.zeros(256, 352, device=dev)
for div in (1, 4, 8, 16, 32):
    rows = ROWS // div
    t = timed(lambda: ops._call("gnb_linear_bwd_weight_tf32", ops._ptr(dz), 256, ops._ptr(h), 336, ops._ptr(dw), 352, rows, 256,
                                336, 0, ops._stream()))
    print(f"rows/{div}: {t:.1f} us -> {t * div:.0f} us per full edge list, {4.0 * rows * 592 / t / 1e6:.2f} TB/s operand bytes", flush=True)
