#!/usr/bin/env python
"""Benchmark of the B200 DynEdge hot path (BASELINE.json metric: DynEdge events/sec, fwd+bwd and inference).

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path (headline workload train512)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the host CPU cores
    python bench.py --workload {infer1024,highmult20k,percentile16,prometheus50,microbench,tito256}   # the other BASELINE configs

One "step" = one pass of the hot path over one batch of synthetic IceCube-like events:
  headline `value`  : training step of BASELINE configs[2] -- device-resident x/batch/n_pulses (no edge_index)
                      -> kNN graph -> DynEdge fwd -> direction(vMF)+energy(LogCosh) heads and loss -> bwd ->
                      (NCCL sum all-reduce of the flat gradient buffer when N > 1, its tail overlapped with the backward of
                      the early layers; 1 / world folded into the optimizer) -> Adam step (one launch on the flat buffers);
                      512 events per GPU (weak scaling), events/s summed over all ranks. Precision mode tf32x3 (split-operand
                      forward GEMMs, single-pass tf32 backward GEMMs: outputs ~1e-6, gradients < 1e-3 from the oracle); the
                      single-pass tf32 mode (gradients 2.3e-3) is reported beside it under `alt_precision`.
  `inference`       : BASELINE configs[1] -- forward + energy head on 1024 events per GPU, no collective; own e2e + roofline.
  `e2e`             : the training step driven from pinned HOST buffers through the public API
                      (H2D copy of the batch and D2H read of the loss inside the timed region).
  `kernels`         : every kernel >= 1 % of the step -- us per step, algorithmic FLOP / bytes, fraction of its roofline.
Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

# arithmetic type of the path's dominant contractions (not a precision claim: config.precision states the tolerance)
DTYPE = {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32", "bf16": "bf16", "bf16x3": "bf16", "mixed16": "f16", "f16": "f16"}
METRIC = "dynedge_train_events_per_sec"
UNIT = "events/s"
POOLS = ["min", "max", "mean", "sum"]
TOLERANCE = {"mixed16": "out 2e-5 / grad 1e-3 vs the fp64 oracle (tests/test_gpu_bf16.py, tests/test_gpu_train_step.py; measured "
                        "9e-6 / 5.5e-4): per-edge tensors as power-of-two scaled fp16 planes, forward 3 products (fp32 grade), "
                        "backward 1 product (tf32 grade); node-level GEMMs as tf32x3",
             "f16": "out 1e-3 / grad 3e-3 (per-edge tensors as ONE power-of-two scaled fp16 plane = tf32's 11-bit significand at half the "
                    "bytes, node-level GEMMs single-pass tf32; measured 6e-4 / 2e-3, tests/test_gpu_bf16.py)",
             "bf16x3": "out 2e-5 / grad 1e-3 (per-edge tensors as two bf16 planes, 3 products forward and backward; measured 9e-6 / 5.5e-4)",
             "bf16": "out 5e-3 / grad 1.5e-2 (north_star's looser mode: per-edge tensors as ONE bf16 plane; measured 2.3e-3 / 5.7e-3, "
                     "tests/test_gpu_bf16.py)",
             "tf32x3": "out 2e-5 / grad 1e-3 vs the fp64 oracle (tests/test_gpu_tf32x3.py)",
             "tf32": "out 1e-3 / grad 3e-3 (single-pass tf32, measured 8e-4 / 2.3e-3, tests/test_gpu_tc.py)",
             "fp32": "out 1e-3 / grad 1e-3 (SIMT fp32, measured 4e-7 / 8e-7)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train512",
                    choices=["train512", "infer1024", "highmult20k", "percentile16", "prometheus50", "microbench", "tito256"])
    ap.add_argument("--events", type=int, default=512, help="training events per GPU (configs[2])")
    ap.add_argument("--infer-events", type=int, default=1024, help="inference events per GPU (configs[1])")
    ap.add_argument("--precision", default=os.environ.get("GNB_PRECISION", "mixed16"),
                    choices=["mixed16", "f16", "bf16x3", "bf16", "tf32x3", "tf32", "fp32"])
    ap.add_argument("--cpu-events", type=int, default=128, help="events of the bounded CPU-baseline sample (BASELINE.md 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--no-alt-precision", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce behind the backward instead of the overlapped pair")
    ap.add_argument("--repeats", type=int, default=3, help="repetitions of the K timed steps; the MEDIAN repetition is reported")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def measure_tf32_peak(dev, sustained_s: float = 2.0):
    """TF32 dense peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with
    allow_tf32, best of 10 with CUDA events (burst) and back to back for `sustained_s` seconds (sustained). cuBLAS is only
    the measuring stick for the roofline denominator; it is not on the product path."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        reps = max(10, int(sustained_s * best * 1e12 / flops))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b, c
        return {"tf32_tflops": round(best, 1), "tf32_tflops_sustained": round(sustained, 1),
                "how": f"torch.matmul fp32 with allow_tf32, 8192^3: best of 10 (burst), {reps} back to back (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def full_peaks(dev, measure=True):
    pk = peaks()
    tf = measure_tf32_peak(dev) if measure else {"tf32_tflops": 0.0, "tf32_tflops_sustained": 0.0, "how": ""}
    pk.update({"tf32_tflops": tf["tf32_tflops"], "tf32_tflops_sustained": tf["tf32_tflops_sustained"], "tf32_how": tf["how"]})
    return pk


# --------------------------------------------------------------------------------------------- #
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------- #
class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md) through NVML.
    `sample()` is called by the timing loop right after a step has been enqueued, i.e. while the GPU is executing
    it, from the main thread: a concurrently polling `nvidia-smi -lms` process or NVML thread was measured to
    stall kernel launches for 30-150 ms now and then, which is not what is being benchmarked."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu_index, self.samples, self.reasons, self.handle, self.smax, self.nvml = gpu_index, [], set(), None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[self.gpu_index].isdigit() else self.gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.handle = None

    def sample(self):
        if self.handle is None:
            return
        try:
            self.samples.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        if self.handle is None:
            return self._nvidia_smi_once()
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml, one sample per timed step while the step executes"}

    def _nvidia_smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(t) for t in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}


# --------------------------------------------------------------------------------------------- #
# workload
# --------------------------------------------------------------------------------------------- #
def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def _global_batch(num_events: int, world: int, seed: int, replicate: bool, **gen):
    from graphnet_b200.synthetic import make_batch
    if world == 1 or not replicate:
        return make_batch(num_events * world, seed=seed, **gen)
    one = make_batch(num_events, seed=seed, **gen)
    sizes = np.tile(one["n_pulses"], world)
    return {"x": np.tile(one["x"], (world, 1)), "n_pulses": sizes, "energy": np.tile(one["energy"], world),
            "direction": np.tile(one["direction"], (world, 1)),
            "batch": np.repeat(np.arange(len(sizes), dtype=np.int64), sizes.astype(np.int64))}


def host_batches(num_events: int, count: int, seed0: int, rank: int = 0, world: int = 1, replicate: bool = False, **gen):
    """`count` pinned host batches. With world > 1 every rank builds the same GLOBAL batch of num_events * world events
    and keeps the events `assign_events` gives it (balanced on the cost model n + beta n^2, largest events spread first,
    SURVEY 8e): the global batch is exactly num_events * world events per step, the per-rank event counts differ slightly.
    `event_index` = the global indices of the rank's events (ascending).
    replicate=True (what the weak-scaling bench uses): the global batch is `world` replicas of the SAME num_events-event
    draw, so the per-GPU work is exactly the 1-GPU work at every N (an independent draw of N x 512 events differs from the
    1-GPU draw by a few percent in pulses -- 79.3 k vs 82.5 k per rank measured -- which would be read as scaling loss)."""
    from graphnet_b200.distributed import assign_events
    out = []
    for i in range(count):
        raw = _global_batch(num_events, world, seed0 + i, replicate, **gen)
        index = np.arange(num_events * world)
        if world > 1:
            index = assign_events(raw["n_pulses"], world)[rank]
            sizes = raw["n_pulses"].astype(np.int64)
            starts = np.concatenate([[0], np.cumsum(sizes)])
            rows = np.concatenate([np.arange(starts[e], starts[e + 1]) for e in index])
            raw = {"x": raw["x"][rows], "batch": np.repeat(np.arange(len(index), dtype=np.int64), sizes[index]),
                   "n_pulses": raw["n_pulses"][index], "energy": raw["energy"][index], "direction": raw["direction"][index]}
        hb = {k: _pin(torch.from_numpy(np.ascontiguousarray(raw[k]))) for k in ("x", "batch", "n_pulses", "energy", "direction")}
        hb["event_index"] = torch.from_numpy(np.ascontiguousarray(index))
        out.append(hb)
    return out


DEVICE_KEYS = ("x", "batch", "n_pulses", "energy", "direction")


def to_device(hb, dev):
    return {k: hb[k].to(dev, non_blocking=True) for k in DEVICE_KEYS if k in hb}


class DeviceStager:
    """Persistent device-side staging buffers for the end-to-end loop: every step copies its pinned host batch
    into the same preallocated device memory (what an input pipeline does), so the H2D copies are inside the timed
    region but no allocator traffic is."""

    def __init__(self, host_batches_list, dev, keys=DEVICE_KEYS):
        self.bufs, self.keys = {}, [k for k in keys if k in host_batches_list[0]]
        for k in self.keys:
            biggest = max(host_batches_list, key=lambda hb: hb[k].shape[0])[k]
            self.bufs[k] = torch.empty_like(biggest, device=dev)

    def load(self, hb):
        out = {}
        for k in self.keys:
            dst = self.bufs[k][: hb[k].shape[0]]
            dst.copy_(hb[k], non_blocking=True)
            out[k] = dst
        return out


class Trainer:
    """The public-API training step: KNNEdges -> DynEdge -> heads/loss -> backward -> all-reduce -> Adam."""

    def __init__(self, dev, world: int, nb_inputs: int = 7, overlap: bool = True):
        from graphnet_b200 import ops as _ops
        from graphnet_b200.distributed import FlatAdam, FlatGradAllReduce
        from graphnet_b200.models.gnn import DynEdge
        from graphnet_b200.models.graphs.edges import KNNEdges
        from graphnet_b200.tasks import FusedEnergyDirectionTask
        torch.manual_seed(0)
        self.backbone = DynEdge(nb_inputs, global_pooling_schemes=POOLS).to(dev)
        self.tasks = FusedEnergyDirectionTask(128).to(dev)      # both heads + both losses in two CUDA kernels
        self.energy, self.direction = self.tasks.energy, self.tasks.direction
        self.edges = KNNEdges(8)
        self.params = list(self.backbone.parameters()) + list(self.energy.parameters()) + \
            list(self.direction.parameters())
        self.reducer = FlatGradAllReduce(self.params)
        _ops.ACCUMULATE_INTO_GRAD = True      # gradients land directly in the flat all-reduce buffer
        self.opt = FlatAdam(self.reducer, lr=1e-3, eps=1e-3)     # torch.optim.Adam semantics, one launch on the flat buffers
        self.world = world
        self.overlap = overlap and world > 1
        n_conv = len(self.backbone._conv_layers)
        self.tail_param, self.tail_layer = 4 * (n_conv - 1), n_conv - 1   # last conv layer onward: final once its backward ran

    def make_data(self, db):
        from graphnet_b200 import Data
        return Data(x=db["x"], batch=db["batch"], n_pulses=db["n_pulses"])

    def train_step(self, db, grad_probe=None):
        # the flat gradient buffer is zero here: allocated zeroed, then zeroed again by every Adam step behind its read
        data = self.edges(self.make_data(db))
        h = self.backbone(data)
        self.last_backbone_out = h.detach()               # (a view, no copy: parity tests read the read-out decisions from it)
        loss, _, _ = self.tasks(h, db["energy"], db["direction"])
        if self.overlap:
            self.reducer.arm_overlap(self.tail_param, self.tail_layer)
        loss.backward()
        scale = self.reducer.all_reduce_sum_overlapped() if self.overlap else self.reducer.all_reduce_sum()
        if grad_probe is not None:
            grad_probe(self.reducer.flat, scale)          # test hook: the reduced gradient right before the optimizer
        self.opt.step(zero_grad=True, grad_scale=scale)
        return loss

    @torch.no_grad()
    def infer_step(self, db):
        data = self.edges(self.make_data(db))
        return self.energy(self.backbone(data))


def timed_loop(fn, batches, steps, warmup, flush, e2e_host=None, dev=None, sampler=None, d2h="scalar"):
    """Per-step CUDA-event timing with an (untimed) L2 flush between steps. Returns seconds; the per-step device times
    are left in `timed_loop.last_step_ms`."""
    # rotate only over batches that the warm-up has already seen: a first-seen shape costs one-off cudaMallocs
    # inside the caching allocator (measured: a 100-150 ms hiccup), which is not steady-state step time
    nb = max(1, min(warmup, len(batches if e2e_host is None else e2e_host)))
    stager = None
    if e2e_host is None:
        batches = batches[:nb]
    else:
        e2e_host = e2e_host[:nb]
        stager = DeviceStager(e2e_host, dev)
    # every rotating shape is visited twice before timing so that PyTorch's caching allocator has converged (a
    # first or second visit of a shape can still trigger a cudaMalloc of a few MB that blocks for up to 100 ms)
    for i in range(max(warmup, 2 * nb)):
        fn(batches[i % len(batches)] if e2e_host is None else stager.load(e2e_host[i % len(e2e_host)]))
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    from graphnet_b200 import ops as _ops
    total_ms, wall, host, launches0, per_step = 0.0, 0.0, 0.0, _ops.kernel_launch_count(), []
    d2h_bytes = 0
    if e2e_host is None:
        # Device-resident inputs: the K steps are enqueued back to back like a training loop does (no host synchronisation
        # between steps -- with one, every step began on an idle GPU and carried ~0.25 ms of host latency before its first
        # launch). Each step is bracketed by its own pair of CUDA events; the L2 flush sits in the stream BETWEEN the end
        # event of step i and the begin event of step i + 1, so it still evicts L2 and is not part of any step's time.
        events = []
        t0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(float(i))                   # 256 MiB write: evicts L2 between timed steps
            beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            beg.record()
            fn(batches[i % len(batches)])
            end.record()
            events.append((beg, end))
            if sampler is not None:
                sampler.sample()                    # GPU is executing the enqueued steps right now
        host = time.perf_counter() - t0             # host time to enqueue the K steps (no sync)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        per_step = [b.elapsed_time(e) for b, e in events]
        total_ms = sum(per_step)
    else:
        # End to end: every step's inputs start in pinned host memory and every step's result is read back to the host, all
        # inside ONE wall-clock region over the K steps -- pipelined the way an input pipeline + training loop are: the H2D
        # copy of step i + 1 runs on a copy stream into the second staging buffer while step i computes, the result of step
        # i is copied to pinned memory behind the step and read by the host one step later (a lagging loss read-out), so
        # the host never drains the GPU between steps. The in-stream L2 flushes are part of the wall-clock time.
        main = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        stagers = [stager, DeviceStager(e2e_host, dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [None, None]
        res_host, res_evt, events = [None, None], [None, None], []

        def stage(i):
            hb = e2e_host[i % len(e2e_host)]
            with torch.cuda.stream(copy_stream):
                if free[i & 1] is not None:
                    copy_stream.wait_event(free[i & 1])      # the step that last read this staging buffer has finished
                dbs = stagers[i & 1].load(hb)
                ready[i & 1].record(copy_stream)
            return dbs

        def read_result(i):
            res_evt[i & 1].synchronize()
            r = res_host[i & 1]
            return float(r.sum()) if r.numel() > 1 and d2h == "scalar" else r

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        staged = stage(0)
        for i in range(steps):
            flush.fill_(float(i))                   # 256 MiB write: evicts L2 between timed steps
            main.wait_event(ready[i & 1])
            beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            beg.record()
            out = fn(staged).detach()
            if d2h == "scalar" and out.numel() > 1:
                out = out.float().sum()
            if res_host[i & 1] is None or res_host[i & 1].shape != out.shape:
                res_host[i & 1] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            res_host[i & 1].copy_(out, non_blocking=True)                     # D2H read of the loss / the predictions
            d2h_bytes = int(out.numel() * out.element_size())
            end.record()
            res_evt[i & 1] = end
            free[i & 1] = end
            events.append((beg, end))
            if i + 1 < steps:
                staged = stage(i + 1)
            if i > 0:
                _ = read_result(i - 1)
            if sampler is not None:
                sampler.sample()
        _ = read_result(steps - 1)
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        per_step = [b.elapsed_time(e) for b, e in events]
        total_ms = sum(per_step)
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    timed_loop.last_launches = _ops.kernel_launch_count() - launches0      # over all `steps` timed steps
    timed_loop.last_host_ms = host / max(steps, 1) * 1e3
    timed_loop.last_step_ms = per_step
    timed_loop.last_d2h = d2h_bytes
    return (wall if e2e_host is not None else total_ms / 1e3)


def repeat_median(repeats, *args, **kwargs):
    """K timed steps, repeated `repeats` times; the MEDIAN repetition is the reported one (BASELINE.md section 4 asks for
    a median; round 1 reported the fastest repetition). Returns (seconds of the median repetition, stats) where stats
    carries the fastest repetition and the median / min / max of all individual step times."""
    runs = []
    for _ in range(max(1, repeats)):
        sec = timed_loop(*args, **kwargs)
        runs.append((sec, timed_loop.last_launches, timed_loop.last_host_ms, list(timed_loop.last_step_ms), timed_loop.last_d2h))
    order = sorted(range(len(runs)), key=lambda i: runs[i][0])
    med = runs[order[len(order) // 2]]
    allsteps = [ms for r in runs for ms in r[3]]
    timed_loop.last_launches, timed_loop.last_host_ms, timed_loop.last_d2h = med[1], med[2], med[4]
    stats = {"best_repetition_s": runs[order[0]][0], "worst_repetition_s": runs[order[-1]][0],
             "step_ms_median": round(float(np.median(allsteps)), 4), "step_ms_min": round(float(np.min(allsteps)), 4),
             "step_ms_max": round(float(np.max(allsteps)), 4), "timed_steps_total": len(allsteps)}
    return med[0], stats


def max_over_ranks(seconds: float, dev) -> float:
    if not dist.is_initialized():
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ranks(value: float, dev):
    if not dist.is_initialized():
        return [value]
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


# --------------------------------------------------------------------------------------------- #
# roofline of the dominant kernels, timed live with CUDA events on the launching stream
# --------------------------------------------------------------------------------------------- #
def dominant_launches(trainer, db):
    """The heaviest launches of the step (profiles/ launch lists), set up on real graph metadata of one batch with random
    operand values, each as a zero-argument callable through the C-ABI:

      dgrad_scatter: gemm_tc_pair_dual_scatter_kernel, backward of an EdgeConv layer: dh = dz W2 (rows = N*9 padded edge
                     slots, K = 256, 336 output channels) with the ReLU-mask + dP/dQ scatter epilogue (dh never stored)
      agg_fwd:       gemm_tc_pair_kernel<false>, forward m = relu(h W2^T + b2) (K = 336, 256 channels) with the k-sum
                     + mask-bit epilogue (m never stored), single-pass tf32
      agg_fwd_x3:    gemm_tc_pair_kernel<true>, the same launch on split operands (3 tf32 products per K step)
    """
    from graphnet_b200 import ops
    dev = db["x"].device
    data = trainer.edges(trainer.make_data(db))
    graph = data.knn_graph()
    n, width = graph.n, graph.width
    assert width == 9
    rows = n * width
    lin = trainer.backbone._conv_layers[1].nn[2]          # Linear(336, 256)
    hid, cout = lin.in_features, lin.out_features
    ntile = (n + 13) // 14
    w2 = lin.weight.detach()
    h_raw = torch.rand(rows, hid, device=dev)
    h = ops._round_pad(h_raw)
    w2p = ops._tc_pack_weight(w2, (0,), (hid,))
    w_hi, w_lo = ops._tc_pack_weight_split(w2, (0,), (hid,))
    b2 = lin.bias.detach()
    y = torch.empty(n, cout, device=dev)
    maskbits = torch.empty(ntile * cout * 4, dtype=torch.int32, device=dev)
    dz = ops._round_pad(torch.randn(rows, cout, device=dev))
    wt = ops._tc_pack_weight(w2.t().contiguous(), (0,), (cout,))
    mld = 4 * ((hid + 127) // 128)
    hmask = torch.randint(-2 ** 31, 2 ** 31 - 1, (ntile * 126, mld), dtype=torch.int32, device=dev)
    dpq = torch.zeros(n, 2 * hid, device=dev)

    def agg_fwd():
        ops._call("gnb_edge_linear_agg_fwd_tf32", ops._ptr(h), hid, hid, ops._ptr(w2p), w2p.shape[1], ops._ptr(b2),
                  ops._ptr(graph.deg), n, cout, 1, ops._ptr(y), cout, ops._ptr(maskbits), ops._stream())

    def agg_fwd_x3():
        ops._call("gnb_edge_linear_agg_fwd_tf32x3", ops._ptr(h_raw), hid, hid, ops._ptr(w_hi), ops._ptr(w_lo), w_hi.shape[1],
                  ops._ptr(b2), ops._ptr(graph.deg), n, cout, ops._ptr(y), cout, ops._ptr(maskbits), ops._stream())

    def dgrad_scatter():
        ops._call("gnb_edge_hidden_dgrad_scatter_tf32", ops._ptr(dz), cout, cout, ops._ptr(wt), wt.shape[1], ops._ptr(hmask),
                  mld, hid, ops._ptr(graph.nbr), n, ops._ptr(dpq), 2 * hid, ops._stream())

    # mixed16: fp16 planes (power-of-two scaled) of the same operands
    def f16_planes(t, two, dst_cols=None, transpose=False):
        r, c = t.shape
        dr, dc = (c, dst_cols or r) if transpose else (r, dst_cols or c)
        p0 = torch.empty(dr, dc, dtype=torch.float16, device=dev)
        p1 = torch.empty(dr, dc, dtype=torch.float16, device=dev) if two else None
        ops._call("gnb_to_f16_planes", ops._ptr(t), c, r, c, ops._ptr(p0), ops._ptr(p1), dc, dc, 1 if transpose else 0, ops._stream())
        return p0, p1
    hw = torch.zeros(1, dtype=torch.int32, device=dev)
    zw = torch.zeros(1, dtype=torch.int32, device=dev)
    ops._call("gnb_absmax_bits", ops._ptr(h_raw), hid, rows, hid, 0, ops._ptr(hw), ops._stream())
    ops._call("gnb_absmax_bits", ops._ptr(dz), cout, rows, cout, 0, ops._ptr(zw), ops._stream())
    hscale = 2.0 ** (14 - (((int(hw.item()) >> 23) & 0xFF) - 127))
    zscale = 2.0 ** (14 - (((int(zw.item()) >> 23) & 0xFF) - 127))
    h16 = f16_planes(h_raw * hscale, True)
    hld64, cld64 = (hid + 63) // 64 * 64, (cout + 63) // 64 * 64
    w16 = f16_planes(w2.contiguous(), True, dst_cols=hld64)
    wt16 = f16_planes(w2.contiguous(), False, dst_cols=cld64, transpose=True)
    dz16 = (dz * zscale).half()
    dwg = torch.zeros(cout, hid, device=dev)
    rowmask = torch.randint(-2 ** 31, 2 ** 31 - 1, (ntile * 126, cout // 32), dtype=torch.int32, device=dev)
    g16 = torch.randn(n, cout, device=dev).half()

    def agg_fwd_f16x3():
        ops._call("gnb_edge_linear_agg_fwd_f16", ops._ptr(h16[0]), ops._ptr(h16[1]), hid, hid, ops._ptr(w16[0]), ops._ptr(w16[1]),
                  hld64, ops._ptr(b2), ops._ptr(graph.deg), n, cout, 0, ops._ptr(y), cout, ops._ptr(maskbits), ops._ptr(hw),
                  ops._stream())

    def agg_fwd_f16x1():
        ops._call("gnb_edge_linear_agg_fwd_f16", ops._ptr(h16[0]), ops._ptr(None), hid, hid, ops._ptr(w16[0]), ops._ptr(None),
                  hld64, ops._ptr(b2), ops._ptr(graph.deg), n, cout, 1, ops._ptr(y), cout, ops._ptr(None), ops._ptr(hw),
                  ops._stream())

    def dgrad_scatter_f16():
        ops._call("gnb_edge_hidden_dgrad_scatter_f16", ops._ptr(dz16), cout, cout, ops._ptr(wt16[0]), cld64, ops._ptr(hmask), mld,
                  hid, ops._ptr(graph.nbr), n, ops._ptr(dpq[:, hid:]), 2 * hid, ops._ptr(dpq), 2 * hid, ops._ptr(None), 0,
                  ops._ptr(zw), ops._stream())

    def wgrad_f16():
        ops._call("gnb_linear_bwd_weight_f16", ops._ptr(dz16), cout, ops._ptr(h16[0]), ops._ptr(None), hid, ops._ptr(dwg), hid, rows,
                  cout, hid, ops._ptr(zw), ops._ptr(hw), ops._stream())

    def wgrad_f16_masked():
        ops._call("gnb_linear_bwd_weight_f16_masked", ops._ptr(g16), ops._ptr(rowmask), ops._ptr(h16[0]), hid, ops._ptr(dwg), hid, n,
                  cout, hid, ops._ptr(zw), ops._ptr(hw), ops._stream())

    def dgrad_scatter_f16_masked():
        ops._call("gnb_edge_hidden_dgrad_scatter_f16_masked", ops._ptr(g16), ops._ptr(rowmask), cout, ops._ptr(wt16[0]), cld64,
                  ops._ptr(hmask), mld, hid, ops._ptr(graph.nbr), n, ops._ptr(dpq[:, hid:]), 2 * hid, ops._ptr(dpq), 2 * hid,
                  ops._ptr(None), 0, ops._ptr(zw), ops._stream())

    # the fused EdgeConv forward of the fp16-plane modes: PQ (hoisted first Linear) + neighbour table in, y + bit masks out,
    # training side outputs = plane 0 of h and the row-major ReLU bits
    pq = torch.randn(n, 2 * hid, device=dev)
    pqw = torch.zeros(1, dtype=torch.int32, device=dev)
    ops._call("gnb_absmax_bits", ops._ptr(pq), 2 * hid, n, 2 * hid, 1, ops._ptr(pqw), ops._stream())
    h0f = torch.empty(rows, hid, dtype=torch.float16, device=dev)
    hbytes = torch.empty(ntile * 126, mld * 4, dtype=torch.uint8, device=dev)

    def fused_fwd_f16(planes=2, side=True):
        ops._call("gnb_edgeconv_fused_fwd_f16", ops._ptr(pq), 2 * hid, hid, ops._ptr(graph.nbr), ops._ptr(graph.deg), n,
                  ops._ptr(w16[0]), ops._ptr(w16[1] if planes == 2 else None), hld64, ops._ptr(b2), cout, 0, ops._ptr(y), cout,
                  ops._ptr(maskbits if side else None), ops._ptr(h0f if side else None), hid, ops._ptr(hbytes if side else None),
                  mld * 4, ops._ptr(pqw), 1, ops._stream())         # (timing: random PQ, the layout the executor uses)

    keep = (h, h_raw, w2p, w_hi, w_lo, b2, y, maskbits, dz, wt, hmask, dpq, graph, hw, zw, h16, w16, wt16, dz16, dwg, rowmask, g16,
            pq, pqw, h0f, hbytes)
    e_real = int(graph.deg.sum().item())
    return {"agg_fwd": agg_fwd, "agg_fwd_x3": agg_fwd_x3, "dgrad_scatter": dgrad_scatter, "agg_fwd_f16x3": agg_fwd_f16x3,
            "dgrad_scatter_f16": dgrad_scatter_f16, "wgrad_f16": wgrad_f16, "wgrad_f16_masked": wgrad_f16_masked,
            "dgrad_scatter_f16_masked": dgrad_scatter_f16_masked, "fused_fwd_f16x3": fused_fwd_f16, "agg_fwd_f16x1": agg_fwd_f16x1,
            "fused_fwd_f16x3_inference": lambda: fused_fwd_f16(2, False), "fused_fwd_f16x1_inference": lambda: fused_fwd_f16(1, False),
            "rows": rows, "n": n,
            "edges": e_real, "hid": hid, "cout": cout, "mld": mld, "keep": keep}


def _time_launch(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg.record()
    for _ in range(reps):
        fn()
    end.record()
    torch.cuda.synchronize()
    return beg.elapsed_time(end) / 1e3 / reps


def _traffic(name):
    """ncu --set full DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the committed captures."""
    for rnd in ("r02", "r01"):
        path = os.path.join(ROOT, "profiles", rnd, "roofline_traffic.json")
        if os.path.exists(path):
            t = json.load(open(path))
            if name in t:
                return dict(t[name], source=f"profiles/{rnd}/roofline_traffic.json")
            if name == "dgrad_scatter" and "dram_bytes_per_launch" in t:
                return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "rows": t.get("rows"), "source": t.get("source")}
    return None


def roofline_top_kernel(trainer, db, pk, precision, inference=False):
    """Dominant kernel of the step in `precision`, timed alone with CUDA events on the launching stream (operands > 0.7 GB
    per launch exceed the 126 MB L2). Algorithmic FLOPs per launch = 2 * E * 336 * 256 with E = the real edge count
    (deg * 2 * in * out per node; the split kernel EXECUTES three tf32 products per algorithmic product and is
    charged the algorithmic count only). peak = the TF32 dense peak measured in this run (burst: the kernel is timed alone).
    Training: the forward aggregating launch in tf32x3 (gemm_tc_pair_kernel<true>, the step's heaviest family) with the
    backward scattering launch beside it; inference / single-pass: the single-pass aggregating launch."""
    if precision == "fp32":
        return roofline_fp32_kernel(trainer, db, pk)
    d = dominant_launches(trainer, db)
    rows, n, e_real, hid, cout = d["rows"], d["n"], d["edges"], d["hid"], d["cout"]
    flops = 2.0 * e_real * hid * cout
    peak = pk["tf32_tflops"]
    bytes_f = 4.0 * rows * hid + 4.0 * n * cout + 16.0 * ((n + 13) // 14) * cout
    bytes_b = 4.0 * rows * cout + 4.0 * rows * d["mld"] + 4.0 * n * 2 * hid

    def entry(kernel, sec, by, executed=1.0, traffic=None, peak=peak):
        ach = flops / sec / 1e12
        return {"kernel": kernel, "launch_ms": round(sec * 1e3, 4), "achieved": round(ach, 3), "unit": "TFLOP/s",
                "frac": round(ach / peak, 5), "peak": peak, "executed_tflops": round(ach * executed, 3),
                "executed_frac": round(ach * executed / peak, 5),
                "hbm_view": {"algorithmic_bytes": by, "achieved_gbs": round(by / sec / 1e9, 1),
                             "frac": round(by / sec / 1e9 / pk["hbm_gbs"], 4)}, "traffic": traffic}
    sec_f = _time_launch(d["agg_fwd"])
    sec_b = _time_launch(d["dgrad_scatter"])
    fwd1 = entry("gemm_tc_pair_kernel<false>, aggregating epilogue (single-pass tf32): m = relu(h W2^T + b2) summed over the "
                 "k slots, 336 -> 256", sec_f, bytes_f, 1.0, _traffic("agg_fwd"))
    bwd = entry("gemm_tc_pair_dual_scatter_kernel: dh = dz W2 (256 -> 336), ReLU mask, dP/dQ reduction over the padded edge "
                "list (single-pass tf32; both 256-channel groups from one resident dz tile)", sec_b, bytes_b, 1.0,
                _traffic("dgrad_scatter"))
    if precision == "mixed16" and not inference:
        # kind::f16 kernels: the denominator is the measured dense bf16 / fp16 burst peak of MEASURED_PEAKS.json.
        # The three per-edge launches of a layer in this mode, none of which reads or writes a [N k, .] fp32 tensor:
        p16 = pk["bf16_tflops"]
        ntile = (n + 13) // 14
        by_f = 4.0 * n * 2 * hid + 4.0 * 10 * n + 4.0 * n * cout + 16.0 * ntile * cout + 2.0 * rows * hid + 4.0 * rows * d["mld"]
        by_b = 2.0 * n * cout + 4.0 * rows * (cout // 32) + 4.0 * rows * d["mld"] + 4.0 * rows + 4.0 * n * 2 * hid
        by_w = 2.0 * rows * hid + 2.0 * n * cout + 4.0 * rows * (cout // 32)
        top = entry("gemm_f16_pair_agg_fused_kernel<2> (tcgen05 cta_group::2 kind::f16 M256xN256xK16; fused EdgeConv forward): "
                    "builder warps gather P_i + Q_j (fp32), ReLU, scale, split into two fp16 planes straight into the swizzled B tile; "
                    "3 MMAs per algorithmic product against the TMA-streamed W2 planes; epilogue = bias + ReLU + k-sum + 126 mask "
                    "bits per (tile, channel); side outputs plane 0 of h + ReLU bits for the backward pass. h [N k, 336] never "
                    "crosses HBM in fp32. Inside the kernel the builder warps (L1 wavefronts, gather latency, issue slots) are the long pole, not the tensor pipe",
                    _time_launch(d["fused_fwd_f16x3"]), by_f, 3.0, _traffic("fused_fwd_f16x3"), p16)
        bwd16 = entry("gemm_f16_pair_scatter_build_kernel (kind::f16): dh = dz W2 (256 -> 336) with dz expanded in shared memory from "
                      "fp16(g) [N, 256] and the row-major ReLU bits (dz never stored); ReLU mask + dP slot sums + dQ fp32 reductions "
                      "in the epilogue (dh never stored). Bound by the epilogue's 213 M L2 reductions",
                      _time_launch(d["dgrad_scatter_f16_masked"]), by_b, 1.0, _traffic("dgrad_scatter_f16_masked"), p16)
        wg = entry("gemm_f16_wgrad_build_kernel (kind::f16, MN-major operands): dW2 = dz^T h with dz expanded in shared memory from "
                   "fp16(g) and the ReLU bits; reads plane 0 of h only", _time_launch(d["wgrad_f16_masked"]), by_w, 1.0,
                   _traffic("wgrad_f16_masked"), p16)
        out = {"bound": "tensor", "achieved": top["achieved"], "peak": p16, "unit": "TFLOP/s", "frac": top["frac"],
               "traffic": top["traffic"], "kernel": top["kernel"], "launch_ms": top["launch_ms"],
               "peak_source": "dense bf16 burst peak of MEASURED_PEAKS.json (kind::f16 runs fp16 and bf16 at the same rate)",
               "executed_tflops": top["executed_tflops"], "executed_frac": top["executed_frac"], "hbm_view": top["hbm_view"],
               "rows": rows, "edges": e_real, "algorithmic_flops_per_launch": flops,
               "backward_launch": bwd16, "weight_gradient_launch": wg}
        return _roofline_by_intensity(out, flops, pk)
    if precision == "f16" and inference:
        p16 = pk["bf16_tflops"]
        top = entry("gemm_f16_pair_agg_fused_kernel<1> (tcgen05 cta_group::2 kind::f16 M256xN256xK16; fused EdgeConv forward on ONE "
                    "power-of-two scaled fp16 plane): builder warps gather P_i + Q_j (fp32), ReLU, scale, fp16 into the swizzled B tile; "
                    "W2's plane resident in shared memory (no weight streaming); epilogue = bias + ReLU + k-sum. Neither h [N k, 336] "
                    "nor m [N k, 256] exists in HBM", _time_launch(d["fused_fwd_f16x1_inference"]),
                    4.0 * n * 2 * hid + 4.0 * 10 * n + 4.0 * n * cout, 1.0, _traffic("fused_fwd_f16x1_inference"), p16)
        return _roofline_by_intensity(
            {"bound": "tensor", "achieved": top["achieved"], "peak": p16, "unit": "TFLOP/s", "frac": top["frac"],
             "traffic": top["traffic"], "kernel": top["kernel"], "launch_ms": top["launch_ms"],
             "peak_source": "dense bf16 burst peak of MEASURED_PEAKS.json (kind::f16 runs fp16 and bf16 at the same rate)",
             "executed_tflops": top["executed_tflops"], "executed_frac": top["executed_frac"], "hbm_view": top["hbm_view"],
             "rows": rows, "edges": e_real, "algorithmic_flops_per_launch": flops}, flops, pk)
    if precision == "tf32x3" and not inference:
        sec_x = _time_launch(d["agg_fwd_x3"])
        top = entry("gemm_tc_pair_kernel<true> (tcgen05 cta_group::2 kind::tf32 M256xN256xK8, split operands: 3 MMAs per K step, "
                    "TMA 4 x 48 KiB stages + in-kernel hi/lo splitter, TMEM double-buffered), aggregating epilogue: "
                    "m = relu(h W2^T + b2) summed over the k slots, 336 -> 256", sec_x, bytes_f, 3.0, _traffic("agg_fwd_x3"))
        others = {"backward_launch": bwd, "single_pass_forward_launch": fwd1}
    elif inference:
        top, others = fwd1, {}
    else:
        top, others = bwd, {"forward_launch": fwd1}
    out = {"bound": "tensor", "achieved": top["achieved"], "peak": peak, "unit": "TFLOP/s", "frac": top["frac"],
           "traffic": top["traffic"], "kernel": top["kernel"], "launch_ms": top["launch_ms"],
           "peak_source": "tf32 dense burst peak measured in this run (" + pk["tf32_how"] + ")",
           "executed_tflops": top["executed_tflops"], "executed_frac": top["executed_frac"], "hbm_view": top["hbm_view"],
           "rows": rows, "edges": e_real, "algorithmic_flops_per_launch": flops}
    out.update(others)
    return _roofline_by_intensity(out, flops, pk)


def _roofline_by_intensity(out, flops, pk):
    """Which roof bounds the launch is decided by the roofline model itself: algorithmic FLOP per algorithmic byte against the
    ridge point (measured dense peak / measured HBM copy peak). Below the ridge the memory roof is the lower one and the line's
    achieved / peak / frac are the HBM figures (the tensor figures stay under `tensor_view`); above it the tensor figures."""
    hv = out.get("hbm_view") or {}
    by = hv.get("algorithmic_bytes")
    if not by:
        return out
    ai, ridge = flops / by, out["peak"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    out["arithmetic_intensity_flop_per_byte"] = round(ai, 1)
    out["ridge_flop_per_byte"] = round(ridge, 1)
    if ai < ridge:
        out["tensor_view"] = {"achieved": out["achieved"], "peak": out["peak"], "unit": out["unit"], "frac": out["frac"],
                              "executed_tflops": out.get("executed_tflops"), "executed_frac": out.get("executed_frac"),
                              "peak_source": out.get("peak_source")}
        out.update({"bound": "hbm", "achieved": hv["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hv["frac"],
                    "peak_source": "HBM copy bandwidth of MEASURED_PEAKS.json (burst: the launch is timed alone)",
                    "bound_rule": "arithmetic intensity below the ridge point: the memory roof is the lower roof"})
    else:
        out["bound_rule"] = "arithmetic intensity above the ridge point: the tensor roof is the lower roof"
    return out


def roofline_fp32_kernel(trainer, db, pk):
    """fp32 precision mode: the SIMT GEMM on the per-edge Linear 336 -> 256."""
    from graphnet_b200 import ops
    data = trainer.edges(trainer.make_data(db))
    graph = data.knn_graph()
    rows = graph.n * graph.width
    e_real = int(graph.deg.sum().item())
    lin = trainer.backbone._conv_layers[1].nn[2]
    h = torch.rand(rows, lin.in_features, device=db["x"].device)
    w, b = lin.weight.detach(), lin.bias.detach()
    sec = _time_launch(lambda: ops.linear_act(h, w, b, ops.ACT_RELU))
    flops = 2.0 * e_real * lin.in_features * lin.out_features
    achieved = flops / sec / 1e12
    peak = pk["tf32_tflops"]
    hbm_bytes = 4.0 * rows * (lin.in_features + lin.out_features)
    return {"bound": "tensor", "kernel": "gemm_f32_kernel<0,0,1> (fp32 SIMT): edge MLP Linear 336->256 + ReLU over the padded edge list",
            "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 5),
            "traffic": None, "peak_source": "tf32 dense burst peak measured in this run", "launch_ms": round(sec * 1e3, 4),
            "rows": rows, "edges": e_real,
            "hbm_view": {"algorithmic_bytes": hbm_bytes, "achieved_gbs": round(hbm_bytes / sec / 1e9, 1),
                         "peak_gbs": pk["hbm_gbs"], "frac": round(hbm_bytes / sec / 1e9 / pk["hbm_gbs"], 4)}}


# --------------------------------------------------------------------------------------------- #
# per-kernel table: device time per kernel family from one CUPTI pass (explanatory only -- every headline number is
# CUDA-event timed without a profiler), algorithmic work from the model's shapes
# --------------------------------------------------------------------------------------------- #
def _family(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    fam = m.group(1) if m else name
    if fam.startswith("gemm_tc_pair_kernel"):
        m2 = re.search(r"gemm_tc_pair_kernel<\s*(\(int\))?\s*(\d)\s*>", name)
        fam = "gemm_tc_pair_kernel<" + {"0": "single", "1": "split", "2": "f16", "3": "f16x3"}.get(m2.group(2) if m2 else "0", "single") + ">"
    return fam


def algorithmic_work(cfg, n, rows, e_real, nseg, precision, train=True):
    """Algorithmic FLOPs (literal GEMM shapes, real edge count) and minimum HBM bytes per kernel family for one step of the
    default-shaped DynEdge described by `cfg` (conv hidden / out widths, post, read-out), following the dispatch rules of
    csrc/gemm_tc.cu::launch_linear (CTA-pair kernel for > 128 output channels and >= 296 row tiles)."""
    work = {}

    def add(fam, flops=0.0, by=0.0, bound="tensor"):
        w = work.setdefault(fam, {"flops": 0.0, "bytes": 0.0, "bound": bound, "launches": 0})
        w["flops"] += flops
        w["bytes"] += by
        w["launches"] += 1

    # 16-bit plane modes (per-edge tensors): planes forward / backward and bytes per stored element
    planes = {"bf16": (1, 1), "bf16x3": (2, 2), "mixed16": (2, 1), "f16": (1, 1)}.get(precision)
    scaled = precision in ("mixed16", "f16")                         # fp16 planes with a power-of-two scale word
    fused_fwd = scaled                                               # csrc/dynedge_exec.cu: ConvBuf::fused
    node_split = precision in ("tf32x3", "bf16x3", "mixed16")       # node-level forward GEMMs on split operands

    def dense_fam(m_rows, n_out, fwd):
        if precision == "fp32":
            return "gemm_f32_kernel"
        if fwd and node_split:
            return "gemm_tc_pair_kernel<split>"
        tiles = (m_rows + 127) // 128
        return "gemm_tc_pair_kernel<single>" if (tiles >= 296 and n_out > 128) else "gemm_tc_linear_kernel"
    tc = precision != "fp32"
    cin = cfg["x0_width"]
    skip_w = cin
    for hid, cout in cfg["conv"]:
        add(dense_fam(n, 2 * hid, True), 2.0 * n * cin * 2 * hid, 4.0 * n * (cin + 2 * hid))
        add("knn_table_split_kernel", 0.0, n * (12.0 + 40.0), "hbm")
        if planes is not None:
            pf, pb = planes
            hb, zb = 2.0 * pf, 2.0 * pb                    # bytes per element of h / dz as stored
            mld = 4 * ((hid + 127) // 128)
            ntile = (n + 13) // 14
            if fused_fwd:
                # one kernel: PQ + neighbour table in, y (+ training: mask words, plane 0 of h, ReLU bits of h) out
                side = (2.0 * rows * hid + 4.0 * rows * mld + 16.0 * ntile * cout) if train else 0.0
                add("gemm_f16_pair_agg_fused_kernel", 2.0 * e_real * hid * cout, 4.0 * n * (2 * hid + cout) + 40.0 * n + side)
            else:
                add("edge_hidden_fwd_node_bf16_kernel", 0.0, hb * rows * hid + 4.0 * n * 2 * hid, "hbm")
                add("gemm_tc_pair_kernel<f16x3>" if pf == 2 else "gemm_tc_pair_kernel<f16>", 2.0 * e_real * hid * cout,
                    hb * rows * hid + 4.0 * n * cout)
            if train and scaled:
                # dz is never stored: both GEMMs expand it from fp16(g) [n, cout] and the row-major ReLU bits
                rowmask = 4.0 * rows * (cout // 32)
                # (the launch also zeroes the Q half of dPQ for the scattering kernel: 4 n hid bytes, csrc/dynedge_exec.cu)
                add("edge_dz_prep_kernel", 0.0, 4.0 * n * cout + 2.0 * n * cout + rowmask + 16.0 * ntile * cout + 4.0 * n * hid, "hbm")
                add("gemm_f16_wgrad_build_kernel", 2.0 * e_real * hid * cout, 2.0 * rows * hid + 2.0 * n * cout + rowmask)
                add("gemm_f16_pair_scatter_build_kernel", 2.0 * e_real * hid * cout,
                    2.0 * n * cout + rowmask + 4.0 * rows * mld + 4.0 * rows + 4.0 * n * 2 * hid)
            elif train:
                add("edge_mask_bwd_bf16_kernel", 0.0, zb * rows * cout + 4.0 * n * cout, "hbm")
                add("gemm_bf_wgrad_kernel", 2.0 * e_real * hid * cout, rows * (zb * cout + hb * hid))
                fam = "gemm_bf_pair_dual_scatter_kernel" if hid > 256 else ("gemm_tc_pair_kernel<f16x3>" if pb == 2 else "gemm_tc_pair_kernel<f16>")
                add(fam, 2.0 * e_real * hid * cout, zb * rows * cout + 4.0 * n * 2 * hid)
                add("zero_block_kernel", 0.0, 4.0 * n * hid, "hbm")
        else:
            add("edge_hidden_fwd_node_kernel", 0.0, 4.0 * rows * hid + 4.0 * n * 2 * hid, "hbm")
            add(dense_fam(rows, cout, True), 2.0 * e_real * hid * cout, 4.0 * rows * hid + 4.0 * n * cout)
        if train:
            if planes is None:
                add("edge_mask_bwd_kernel", 0.0, 4.0 * rows * cout + 4.0 * n * cout, "hbm")
                add("gemm_tc_wgrad_kernel" if tc else "gemm_f32_kernel", 2.0 * e_real * hid * cout, 4.0 * rows * (hid + cout))
                fam = "gemm_tc_pair_dual_scatter_kernel" if (tc and hid > 256) else dense_fam(rows, hid, False)
                add(fam, 2.0 * e_real * hid * cout, 4.0 * rows * cout + 4.0 * n * 2 * hid)
                add("zero_block_kernel", 0.0, 4.0 * n * hid, "hbm")
            add("gemm_tc_wgrad_kernel" if tc else "gemm_f32_kernel", 2.0 * n * cin * 2 * hid, 4.0 * n * (cin + 2 * hid))
            if cin != cfg["x0_width"]:
                add(dense_fam(n, cin, False), 2.0 * n * cin * 2 * hid, 4.0 * n * (2 * cin + 2 * hid))
        cin = cout
        skip_w += cout
    k_in = skip_w
    for j, width in enumerate(cfg["post"]):
        add(dense_fam(n, width, True), 2.0 * n * k_in * width, 4.0 * n * (k_in + width))
        if train:
            add("act_bwd_colsum_kernel", 0.0, 12.0 * n * width, "hbm")
            add("gemm_tc_wgrad_kernel" if tc else "gemm_f32_kernel", 2.0 * n * k_in * width, 4.0 * n * (k_in + width))
            dg_in = k_in - cfg["x0_width"] if j == 0 else k_in
            add(dense_fam(n, 256, False), 2.0 * n * dg_in * width, 4.0 * n * (dg_in + width))
        k_in = width
    pooled = len(POOLS) * k_in
    add("segment_pool_fwd_kernel", 0.0, 4.0 * n * k_in + 8.0 * nseg * pooled, "hbm")
    add("global_vars_kernel", 0.0, 4.0 * n * (cfg["nb_inputs"] + 9 + cfg["x0_ld"]), "hbm")
    if train:
        add("segment_pool_bwd_kernel", 0.0, 4.0 * n * k_in + 8.0 * nseg * pooled, "hbm")
    return work


def kernel_table(step_fn, db, cfg, n, rows, e_real, nseg, precision, pk, train=True, steps=2):
    try:
        from torch.profiler import ProfilerActivity, profile
        step_fn(db)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                step_fn(db)
            torch.cuda.synchronize()
        fam_us, fam_n = {}, {}
        for ev in prof.key_averages():
            us = float(getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0) or 0.0)
            if us <= 0.0:
                continue
            fam = _family(ev.key)
            fam_us[fam] = fam_us.get(fam, 0.0) + us / steps
            fam_n[fam] = fam_n.get(fam, 0) + ev.count / steps
    except Exception as exc:            # CUPTI unavailable: the table is explanatory, the bench line stays valid without it
        return {"unavailable": f"{type(exc).__name__}: {exc}"}
    total = sum(fam_us.values())
    if total <= 0.0:
        return {"unavailable": "the profiler returned no device activity"}
    work = algorithmic_work(cfg, n, rows, e_real, nseg, precision, train)
    rows_out = []
    alg_flops_total = sum(w["flops"] for w in work.values())
    for fam, us in sorted(fam_us.items(), key=lambda kv: -kv[1]):
        if us < 0.01 * total:
            continue
        w = work.get(fam)
        row = {"kernel": fam, "us_per_step": round(us, 1), "share": round(us / total, 4), "launches_per_step": round(fam_n[fam], 1)}
        if w is not None:
            if w["bound"] == "tensor":
                ach = w["flops"] / (us * 1e-6) / 1e12
                f16 = fam.startswith(("gemm_bf_", "gemm_f16_")) or fam.endswith(("<f16>", "<f16x3>"))      # kind::f16 kernels: bf16 / fp16 dense peak
                row.update({"bound": "tensor", "algorithmic_gflop": round(w["flops"] / 1e9, 2), "achieved_tflops": round(ach, 1),
                            "frac": round(ach / (pk["bf16_tflops_sustained"] if f16 else pk["tf32_tflops_sustained"]), 4),
                            "peak": "bf16_tflops_sustained" if f16 else "tf32_tflops_sustained",
                            "hbm_frac": round(w["bytes"] / (us * 1e-6) / 1e9 / pk["hbm_gbs"], 4)})
            else:
                ach = w["bytes"] / (us * 1e-6) / 1e9
                row.update({"bound": "hbm", "algorithmic_mb": round(w["bytes"] / 1e6, 2), "achieved_gbs": round(ach, 1),
                            "frac": round(ach / pk["hbm_gbs"], 4)})
        rows_out.append(row)
    return {"source": f"one CUPTI pass over {steps} steps after the timed loops (explanatory; headline numbers are CUDA-event "
                      "timed without a profiler); tensor fractions against the SUSTAINED dense peak of the kernel's operand type (tf32: "
                      "measured in this run; kind::f16 kernels: MEASURED_PEAKS.json bf16_tflops_sustained), HBM fractions against "
                      "MEASURED_PEAKS.json",
            "device_us_per_step": round(total, 1), "algorithmic_gflop_per_step": round(alg_flops_total / 1e9, 1),
            "kernels": rows_out}


def model_shape(trainer):
    bb = trainer.backbone
    conv = [(c.nn[0].out_features, c.nn[2].out_features) for c in bb._conv_layers]
    post = [m.out_features for m in bb._post_processing if isinstance(m, torch.nn.Linear)]
    nbi = bb._nb_inputs
    x0w = nbi + nbi + 5
    return {"conv": conv, "post": post, "nb_inputs": nbi, "x0_width": x0w, "x0_ld": (x0w + 31) // 32 * 32}


def literal_flops(cfg, n, e_real, nseg, train):
    """the reference's literal algorithm -- per edge 2 (2 C_in H + H C_out), per node the post-processing, per
    event the read-out; backward = 2 x forward GEMM FLOPs minus the first layer's dX."""
    cin = cfg["x0_width"]
    fwd, first_dx = 0.0, 0.0
    for i, (hid, cout) in enumerate(cfg["conv"]):
        f = 2.0 * e_real * (2 * cin * hid + hid * cout)
        fwd += f
        if i == 0:
            first_dx = 2.0 * e_real * 2 * cin * hid
        cin = cout
    k_in = cfg["x0_width"] + sum(c for _, c in cfg["conv"])
    for width in cfg["post"]:
        fwd += 2.0 * n * k_in * width
        k_in = width
    fwd += 2.0 * nseg * len(POOLS) * k_in * 128
    return fwd + (2.0 * fwd - first_dx if train else 0.0)


# --------------------------------------------------------------------------------------------- #
# CPU baseline / reference arm: the oracle (pure-torch restatement of the reference algorithm)
# --------------------------------------------------------------------------------------------- #
def cpu_reference_events_per_sec(num_events: int, steps: int, warmup: int, train: bool = True):
    from types import SimpleNamespace
    from graphnet_b200.synthetic import make_batch
    from graphnet_b200.tasks import DirectionReconstructionWithKappa, EnergyReconstruction
    from oracle.dynedge_oracle import DynEdgeRef, knn_graph_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = DynEdgeRef(7, global_pooling_schemes=POOLS)
    energy, direction = EnergyReconstruction(128), DirectionReconstructionWithKappa(128)
    params = list(model.parameters()) + list(energy.parameters()) + list(direction.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, eps=1e-3)
    raw = make_batch(num_events, seed=20240607)
    x, batch, n_pulses = (torch.from_numpy(raw[k]) for k in ("x", "batch", "n_pulses"))
    e, d = torch.from_numpy(raw["energy"]), torch.from_numpy(raw["direction"])

    def step():
        ei = knn_graph_ref(x[:, :3], 8, batch=batch)
        data = SimpleNamespace(x=x, edge_index=ei, batch=batch, n_pulses=n_pulses)
        if train:
            opt.zero_grad()
            h = model(data)
            loss = energy.compute_loss(energy(h), e) + direction.compute_loss(direction(h), d)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                energy(model(data))

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    return num_events / sec, sec, cores, int(x.shape[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    evs, sec, cores, nodes = cpu_reference_events_per_sec(args.cpu_events, steps, warm, train=True)
    sample = (f"{args.cpu_events} synthetic events ({nodes} pulses), {warm} warm-up + {steps} timed training steps (median) of "
              "the pure-torch oracle (reference's own PyG stack is not installable here)")
    line = {"impl": "reference", "metric": METRIC, "value": round(evs, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": round(evs, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(evs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(args, world):
    return {"workload": "BASELINE configs[2]: DynEdge (nb_inputs=7, k=8, pooling min/max/mean/sum) direction(vMF)+energy"
                        "(LogCosh) training step fwd+bwd+Adam, 512 events/GPU, synthetic IceCube86 pulse maps "
                        "(lognormal pulses/event, median 100, max 5000); configs[1] inference B=1024 under 'inference'",
            "events_per_gpu": args.events, "global_events": args.events * world, "parallelism": f"dp{world}",
            "sharding": "global batch = N replicas of the 512-event draw (identical per-GPU work at every N), spread over the ranks "
                        "by assign_events (cost n + 1.1e-4 n^2 per event, largest first); no data-path collective",
            "precision": f"{args.precision}: {TOLERANCE[args.precision]}",
            "inputs": "x/batch/n_pulses resident in HBM; kNN graph built inside the step",
            "l2": "256 MiB buffer rewritten between timed steps; min(4, warmup) rotating batches",
            "timing": "value: the K steps enqueued back to back (no host synchronisation between steps, as in a training loop), "
                      "each step between its own pair of CUDA events, the L2 flush in the stream between two steps and outside "
                      "both; barrier + synchronize before the first and after the last step. e2e: wall clock over the K steps, "
                      "pipelined like an input pipeline + training loop (H2D of step i + 1 on a copy stream into a second staging "
                      "buffer under step i; the loss of step i copied to pinned memory behind the step and read by the host one "
                      "step later); every H2D, D2H and L2 flush inside the wall-clock region",
            "warmup_executed": "max(W, 2 x rotating batches) untimed steps per timed loop",
            "repeats": f"{args.repeats} repetitions of the K timed steps, the MEDIAN repetition reported (fastest and per-step "
                       "median under 'timing')"}


def _emit(line):
    """Write the result line to the process's ORIGINAL stdout (see `_quiet_stdout`)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    """Point fd 1 at stderr for the rest of the run: native libraries (NCCL's version banner, for one) print to
    stdout, and the driver expects exactly one JSON line there."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def graph_stats(trainer, db):
    data = trainer.edges(trainer.make_data(db))
    g = data.knn_graph()
    return int(g.n), int(g.n * g.width), int(g.deg.sum().item()), int(db["n_pulses"].numel())


def run_train512(args, dev, world, rank, local):
    from graphnet_b200 import ops
    ops.set_precision(args.precision)
    pk = full_peaks(dev, measure=rank == 0)
    trainer = Trainer(dev, world, overlap=not args.no_overlap)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    train_host = host_batches(args.events, 4, seed0=20240607, rank=rank, world=world, replicate=True)
    train_dev = [to_device(hb, dev) for hb in train_host]
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sec_local, tstats = repeat_median(args.repeats, trainer.train_step, train_dev, args.steps, args.warmup, flush,
                                      sampler=sampler if rank == 0 else None)
    launches = timed_loop.last_launches        # kernels of libgraphnet_b200.so launched inside the timed steps
    host_ms = timed_loop.last_host_ms
    rank_ms = gather_ranks(tstats["step_ms_median"], dev)
    sec = max_over_ranks(sec_local, dev)
    events_total = float(args.events * world * args.steps)      # every step processes the whole global batch
    value = events_total / sec
    timing = dict(tstats, value_best_repetition=round(events_total / max_over_ranks(tstats["best_repetition_s"], dev), 2),
                  per_rank_step_ms_median=[round(v, 4) for v in rank_ms],
                  per_rank_spread=round((max(rank_ms) - min(rank_ms)) / max(float(np.mean(rank_ms)), 1e-9), 4))

    # end-to-end through the public API from pinned host buffers
    sec_e2e, _ = repeat_median(args.repeats, trainer.train_step, None, args.steps, args.warmup, flush, e2e_host=train_host, dev=dev)
    sec_e2e = max_over_ranks(sec_e2e, dev)
    h2d = sum(int(train_host[0][k].numel() * train_host[0][k].element_size()) for k in DEVICE_KEYS)
    e2e = {"value": round(events_total / sec_e2e, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    clocks = sampler.stop() if rank == 0 else None

    alt = None
    if not args.no_alt_precision and args.precision in ("tf32x3", "mixed16"):
        alt = []          # the same step in the other precision modes (bf16 = north_star's looser, stated mode)
        for mode in (["bf16", "tf32x3", "tf32"] if args.precision == "mixed16" else ["tf32"]):
            ops.set_precision(mode)
            sec_a, st_a = repeat_median(args.repeats, trainer.train_step, train_dev, args.steps, args.warmup, flush)
            sec_a = max_over_ranks(sec_a, dev)
            alt.append({"precision": f"{mode}: " + TOLERANCE[mode], "value": round(events_total / sec_a, 2), "unit": UNIT,
                        "ms_per_step": round(sec_a / args.steps * 1e3, 3), "step_ms_median": st_a["step_ms_median"]})
        ops.set_precision(args.precision)

    inference = None
    if not args.no_inference:
        inference = run_inference(args, trainer, dev, world, rank, flush, pk)

    roof, table, whole = None, None, None
    if rank == 0:
        roof = roofline_top_kernel(trainer, train_dev[0], pk, args.precision)
        n, rows, e_real, nseg = graph_stats(trainer, train_dev[0])
        cfg = model_shape(trainer)
        if not args.no_kernel_table and world == 1:      # (its steps contain collectives: one rank cannot run them alone)
            table = kernel_table(trainer.train_step, train_dev[0], cfg, n, rows, e_real, nseg, args.precision, pk, train=True)
        lit = literal_flops(cfg, n, e_real, nseg, True)
        alg = sum(w["flops"] for w in algorithmic_work(cfg, n, rows, e_real, nseg, args.precision, True).values())
        step_s = tstats["step_ms_median"] * 1e-3
        whole = {"literal_gflop_per_step": round(lit / 1e9, 1), "hoisted_gflop_per_step": round(alg / 1e9, 1),
                 "literal_tflops": round(lit / step_s / 1e12, 1), "hoisted_tflops": round(alg / step_s / 1e12, 1),
                 "frac_literal_vs_tf32_sustained": round(lit / step_s / 1e12 / max(pk["tf32_tflops_sustained"], 1e-9), 4),
                 "frac_hoisted_vs_tf32_sustained": round(alg / step_s / 1e12 / max(pk["tf32_tflops_sustained"], 1e-9), 4),
                 "note": "literal = the reference's per-edge MLP (SURVEY 8d); hoisted = the GEMMs this path runs (first Linear of "
                         "every EdgeConv moved from edges to nodes), algorithmic count (a split product counted once)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        evs, csec, cores, nodes = cpu_reference_events_per_sec(args.cpu_events, 2, 1, train=True)
        cpu = {"value": round(evs, 3), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_events} of the same synthetic events ({nodes} pulses), 1 warm-up + 2 timed "
                         f"training steps of the pure-torch oracle on {cores} host threads"}
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(sec / args.steps * 1e3, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.precision],
                "data": "synthetic", "config": workload_config(args, world), "e2e": e2e, "gpu_launches": int(launches),
                "gpu_launches_per_step": int(launches) // max(args.steps, 1),
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "inference": inference, "alt_precision": alt,
                "timing": timing, "whole_step": whole, "kernels": table,
                "peaks": {"hbm_gbs": pk["hbm_gbs"], "tf32_tflops": pk["tf32_tflops"],
                          "tf32_tflops_sustained": pk["tf32_tflops_sustained"], "bf16_tflops": pk["bf16_tflops"],
                          "source": pk["source"] + " (MEASURED_PEAKS.json) + tf32 measured in this run"},
                "nodes_per_step_rank0": int(train_host[0]["x"].shape[0]),
                "host_enqueue_ms_per_step": round(host_ms, 3)}
        _emit(line)


def run_inference(args, trainer, dev, world, rank, flush, pk):
    """BASELINE configs[1]: energy-regression inference, 1024 events per GPU, in the single-pass tf32 mode (predictions within
    rel 1e-3: tests/test_gpu_tc.py) and, beside it, in the mode of the training run. Device-resident `value`, `e2e` from
    pinned host buffers (H2D of the batch, D2H of the [B, 1] predictions inside the timed region), own roofline."""
    from graphnet_b200 import ops
    inf_host = host_batches(args.infer_events, 2, seed0=777, rank=rank, world=world, replicate=True)
    inf_dev = [to_device(hb, dev) for hb in inf_host]
    events_total = float(args.infer_events * world * args.steps)
    keep = ops.PRECISION
    out = {"workload": "BASELINE configs[1]: energy-regression inference, 1024 events/GPU, no collective", "unit": UNIT}
    try:
        # headline mode of the inference leg: "f16" (tf32 grade: out 1e-3) when the run's training mode is a tensor-core mode
        head = "f16" if keep in ("mixed16", "f16", "tf32x3", "tf32", "bf16", "bf16x3") else keep
        modes = [head] + [m for m in ("tf32", "bf16", keep) if m != head and keep in ("tf32x3", "mixed16")]
        for mode in dict.fromkeys(modes):
            ops.set_precision(mode)
            sec, st = repeat_median(args.repeats, trainer.infer_step, inf_dev, args.steps, args.warmup, flush)
            sec = max_over_ranks(sec, dev)
            entry = {"precision": f"{mode}: {TOLERANCE[mode].split(' / ')[0]}", "value": round(events_total / sec, 2),
                     "ms_per_step": round(sec / args.steps * 1e3, 3), "step_ms_median": st["step_ms_median"]}
            if mode == head:
                sec_e, _ = repeat_median(args.repeats, trainer.infer_step, None, args.steps, args.warmup, flush, e2e_host=inf_host,
                                         dev=dev, d2h="tensor")
                sec_e = max_over_ranks(sec_e, dev)
                h2d = sum(int(inf_host[0][k].numel() * inf_host[0][k].element_size()) for k in ("x", "batch", "n_pulses"))
                entry["e2e"] = {"value": round(events_total / sec_e, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                                "d2h_bytes_per_step": int(timed_loop.last_d2h)}
                out.update(entry)
                if rank == 0:
                    out["roofline"] = roofline_top_kernel(trainer, inf_dev[0], pk, head, inference=True)
                    if not args.no_kernel_table:
                        n, rows, e_real, nseg = graph_stats(trainer, inf_dev[0])
                        out["kernels"] = kernel_table(trainer.infer_step, inf_dev[0], model_shape(trainer), n, rows, e_real, nseg,
                                                      head, pk, train=False)
            else:
                out.setdefault("alt_precision", []).append(entry)
    finally:
        ops.set_precision(keep)
    return out


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as entry
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "graphnet_b200", "csrc", "libgraphnet_b200.so")):
        entry.build()
    if world > 1:
        dist.barrier()
    if args.workload == "train512":
        run_train512(args, dev, world, rank, local)
    elif args.workload == "infer1024":
        from graphnet_b200 import ops
        ops.set_precision(args.precision)
        pk = full_peaks(dev)
        trainer = Trainer(dev, world)
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        inf = run_inference(args, trainer, dev, world, rank, flush, pk)
        if rank == 0:
            _emit({"metric": "dynedge_inference_events_per_sec", "value": inf["value"], "unit": UNIT, "n_gpus": world,
                   "steps": args.steps, "warmup": args.warmup, "ms_per_step": inf["ms_per_step"], "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                   "config": {"workload": inf["workload"], "precision": inf.get("precision")}, "e2e": inf.get("e2e"), "roofline": inf.get("roofline"),
                   "kernels": inf.get("kernels"), "alt_precision": inf.get("alt_precision")})
    else:
        import bench_workloads
        bench_workloads.run(args, dev, world, rank)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
