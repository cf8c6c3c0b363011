"""Data-parallel plumbing for the DynEdge path: event sharding and the one gradient exchange.

Events are independent graphs (kNN, EdgeConv and pooling never cross `batch` boundaries), so the path
shards by events with no data-path collective; training adds a single sum-all-reduce of one flat fp32
gradient buffer (~5.5 MB for the default DynEdge) divided by the world size -- the mean semantics of
the reference's `Trainer(strategy="ddp")` (src/graphnet/models/easy_model.py:90-110).
"""

from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# Cost model of one event with n pulses on the B200 path: every per-pulse / per-edge kernel is linear in n, the kNN
# builds (4 per forward) are quadratic. Measured at 512 events / GPU (profiles/r02): 7.8 ms for 84 k pulses = 93 ns per
# pulse, 4 x 100 us of kNN for sum n^2 = 3.8e7 pairs = 10.6 ps per pair  ->  beta = 1.1e-4 pulses per pair.
KNN_COST_BETA = 1.1e-4


def event_cost(sizes: Sequence[int], beta: float = KNN_COST_BETA) -> np.ndarray:
    n = np.asarray(sizes, dtype=np.float64)
    return n + beta * n * n


def shard_events(sizes: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous event ranges [lo, hi) per rank, balanced by pulse count (prefix-sum cuts); every rank gets at least
    one event (ValueError when there are fewer events than ranks: an empty shard would divide by zero in the loss)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    if len(sizes) < world_size:
        raise ValueError(f"shard_events: {len(sizes)} events cannot be spread over {world_size} ranks")
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        idx = int(np.searchsorted(csum, target, side="left"))
        idx = max(idx, cuts[-1] + 1)                          # at least one event for rank r - 1 ...
        idx = min(idx, len(sizes) - (world_size - r))         # ... and for every rank still to come
        cuts.append(idx)
    cuts.append(len(sizes))
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def assign_events(sizes: Sequence[int], world_size: int, beta: float = KNN_COST_BETA) -> List[np.ndarray]:
    """Event indices per rank (ascending within a rank), balanced on the cost model `n + beta n^2` with the largest events
    placed first (longest-processing-time greedy): unlike contiguous pulse-balanced ranges this also spreads the few
    multi-thousand-pulse events, whose kNN cost is quadratic and whose one-CTA-per-event kernels set the tail of a step."""
    sizes = np.asarray(sizes, dtype=np.int64)
    if len(sizes) < world_size:
        raise ValueError(f"assign_events: {len(sizes)} events cannot be spread over {world_size} ranks")
    cost = event_cost(sizes, beta)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world_size)
    count = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(sizes), dtype=np.int64)
    remaining = len(sizes)
    for e in order:
        empty = np.flatnonzero(count == 0)
        # keep every rank non-empty: once only as many events remain as there are empty ranks, they go to those ranks
        r = int(empty[np.argmin(load[empty])]) if len(empty) and remaining <= len(empty) else int(np.argmin(load))
        owner[e] = r
        load[r] += cost[e]
        count[r] += 1
        remaining -= 1
    return [np.flatnonzero(owner == r) for r in range(world_size)]


class FlatGradAllReduce:
    """All parameters' gradients live in ONE flat fp32 buffer that is all-reduced with one call (or two: see
    `all_reduce_sum_overlapped`)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = []
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)   # grads accumulate straight into the buffer
            self.offsets.append(off)
            off += p.numel()
        self._side: Optional[torch.cuda.Stream] = None
        self._event: Optional[torch.cuda.Event] = None

    @staticmethod
    def _distributed() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    @property
    def world(self) -> int:
        return dist.get_world_size() if self._distributed() else 1

    def sync_params(self, src: int = 0) -> None:
        """Broadcast rank `src`'s parameters, as DDP does at construction: replicas that were seeded differently would
        otherwise diverge silently."""
        if not self._distributed():
            return
        with torch.no_grad():
            for p in self.params:
                dist.broadcast(p.data, src=src)

    def zero(self) -> None:
        self.flat.zero_()

    def rebind(self) -> None:
        """Re-attach views if an optimizer/zero_grad(set_to_none=True) dropped them."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def all_reduce_mean(self, async_op: bool = False):
        """SUM all-reduce of the flat buffer preceded by the division by the world size (a separate pass; `FlatAdam.step`
        with `grad_scale = 1 / world` after `all_reduce_sum*` folds that division into the optimizer launch instead)."""
        if not self._distributed():
            return None
        self.rebind()
        self.flat.div_(dist.get_world_size())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)

    def all_reduce_sum(self) -> float:
        """One SUM all-reduce on the current stream; returns the `grad_scale` the optimizer must apply (1 / world)."""
        if not self._distributed():
            return 1.0
        self.rebind()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()

    # -- overlapped form: the tail of the buffer goes out while the backward of the early layers still runs -------------
    def arm_overlap(self, first_tail_param: int, after_conv_layer: int) -> None:
        """Before `loss.backward()`: ask the executor to record an event once the backward of conv layer `after_conv_layer`
        is enqueued; parameters `first_tail_param ...` (that layer's, the post-processing's, the read-out's and, because the
        heads run first in the backward, the task heads') are final from then on."""
        if not self._distributed():
            return
        from . import _lib
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._event = torch.cuda.Event()
            self._event.record()                               # materialises the underlying cudaEvent_t
        self._tail_off = self.offsets[first_tail_param]
        _lib.check(_lib.load().gnb_dynedge_set_backward_event(self._event.cuda_event, int(after_conv_layer)),
                   "gnb_dynedge_set_backward_event")

    def all_reduce_sum_overlapped(self) -> float:
        """After `loss.backward()` returned (everything is enqueued): all-reduce the tail slice on the side stream as soon
        as the armed event fires, the head slice on the current stream behind the backward, then join. Returns grad_scale."""
        if not self._distributed():
            return 1.0
        if self._side is None:
            return self.all_reduce_sum()
        self.rebind()
        main = torch.cuda.current_stream()
        self._side.wait_event(self._event)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.flat[self._tail_off:], op=dist.ReduceOp.SUM)
        if self._tail_off > 0:
            dist.all_reduce(self.flat[:self._tail_off], op=dist.ReduceOp.SUM)
        main.wait_stream(self._side)
        return 1.0 / dist.get_world_size()


class FlatAdam:
    """`torch.optim.Adam` (no amsgrad) over ONE flat fp32 parameter buffer, one CUDA launch per step.

    The reference trains with `optimizer_class=Adam, optimizer_kwargs={"lr": 1e-3, "eps": 1e-3}`
    (examples/04_training/01_train_dynedge.py:128-129, instantiated at models/easy_model.py:215-219). Here every
    parameter becomes a view into `flat_p` (same order as `FlatGradAllReduce.flat`, which holds the gradients), the
    moments live in two more flat buffers and `csrc/optim.cu::gnb_adam_flat` updates all 1.38 M values in one launch
    (torch's fused multi-tensor Adam needs ~80 us for the 26 small tensors of DynEdge; this takes < 10 us) and can
    zero the gradient buffer behind the read. CUDA only: there is no CPU fallback. Under torch.distributed the initial
    parameters are broadcast from rank 0 (what DDP does at construction).
    """

    def __init__(self, reducer: FlatGradAllReduce, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0):
        if not reducer.flat.is_cuda:
            raise RuntimeError("graphnet_b200.FlatAdam: parameters must live on a CUDA device (no CPU fallback)")
        self.reducer = reducer
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.flat_p = torch.empty_like(reducer.flat)
        off = 0
        with torch.no_grad():
            for p in reducer.params:
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                     # the module's parameters now alias the flat buffer
                off += p.numel()
        if reducer._distributed():
            dist.broadcast(self.flat_p, src=0)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_count = 0

    def step(self, zero_grad: bool = False, grad_scale: float = 1.0) -> None:
        from . import ops
        self.reducer.rebind()
        self.step_count += 1
        b1, b2 = self.betas
        step_size = self.lr / (1.0 - b1 ** self.step_count)
        inv_sqrt_bc2 = 1.0 / float(np.sqrt(1.0 - b2 ** self.step_count))
        n = self.flat_p.numel()
        ops._call("gnb_adam_flat", ops._ptr(self.flat_p), ops._ptr(self.reducer.flat), ops._ptr(self.exp_avg),
                  ops._ptr(self.exp_avg_sq), n, step_size, b1, b2, self.eps, inv_sqrt_bc2, self.weight_decay,
                  float(grad_scale), 1 if zero_grad else 0, ops._stream())
