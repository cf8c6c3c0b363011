"""world_size-2 gloo test of the multi-process plumbing (sharding + flat gradient all-reduce)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphnet_b200.distributed import FlatGradAllReduce, shard_events


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    reducer = FlatGradAllReduce(model.parameters())
    sizes = np.array([3, 9, 4, 20, 5, 7, 11, 2])
    lo, hi = shard_events(sizes, world)[rank]
    rng = np.random.default_rng(1)
    feats = [torch.from_numpy(rng.normal(size=(int(n), 6)).astype(np.float32)) for n in sizes]
    reducer.zero()
    loss = sum(model(f).sum() for f in feats[lo:hi]) / max(hi - lo, 1)
    loss.backward()
    reducer.all_reduce_mean()
    torch.save({"flat": reducer.flat.clone(), "range": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])
    assert outs[0]["range"][1] == outs[1]["range"][0]
    # single-process expectation: mean over ranks of each rank's mean-over-its-events gradient
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    sizes = np.array([3, 9, 4, 20, 5, 7, 11, 2])
    rng = np.random.default_rng(1)
    feats = [torch.from_numpy(rng.normal(size=(int(n), 6)).astype(np.float32)) for n in sizes]
    total = 0.0
    for lo, hi in shard_events(sizes, world):
        total = total + sum(model(f).sum() for f in feats[lo:hi]) / max(hi - lo, 1) / world
    total.backward()
    expect = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(outs[0]["flat"], expect, rtol=1e-5, atol=1e-6)


def test_bench_global_batch_sharding_covers_every_event_once():
    """bench.host_batches(world > 1): every rank keeps the events `assign_events` gives it out of ONE global batch -- together
    the shards hold every event exactly once (pulses, labels), balanced on the cost model n + beta n^2, with `batch`
    renumbered from 0 on every rank."""
    import os
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from graphnet_b200.distributed import event_cost
    world, per_gpu = 3, 16
    full = bench.host_batches(per_gpu * world, 1, seed0=5)[0]
    shards = [bench.host_batches(per_gpu, 1, seed0=5, rank=r, world=world)[0] for r in range(world)]
    index = torch.cat([s["event_index"] for s in shards])
    assert sorted(index.tolist()) == list(range(per_gpu * world))                 # every event exactly once
    starts = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(full["n_pulses"].long(), 0)])
    for s in shards:
        idx = s["event_index"]
        assert torch.equal(s["n_pulses"], full["n_pulses"][idx]) and torch.equal(s["energy"], full["energy"][idx])
        assert torch.equal(s["direction"], full["direction"][idx])
        rows = torch.cat([torch.arange(int(starts[e]), int(starts[e + 1])) for e in idx.tolist()])
        assert torch.equal(s["x"], full["x"][rows])
        b = s["batch"]
        assert int(b.min()) == 0 and int(b.max()) == s["n_pulses"].numel() - 1
        assert torch.equal(torch.bincount(b), s["n_pulses"].long())
    cost = [float(event_cost(s["n_pulses"].numpy()).sum()) for s in shards]
    assert max(cost) - min(cost) <= float(event_cost(full["n_pulses"].numpy()).max())   # balanced to within one event


def test_assign_events_cost_balance_and_no_empty_rank():
    import numpy as np
    import pytest
    from graphnet_b200.distributed import assign_events, event_cost, shard_events
    from graphnet_b200.synthetic import make_batch
    sizes = make_batch(4096, seed=2)["n_pulses"]
    for world in (2, 4, 8):
        parts = assign_events(sizes, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(sizes)))
        assert all(np.all(np.diff(p) > 0) for p in parts)                             # ascending within a rank
        cost = event_cost(sizes)
        loads = [cost[p].sum() for p in parts]
        assert (max(loads) - min(loads)) / np.mean(loads) < 1e-3
        contiguous = [cost[lo:hi].sum() for lo, hi in shard_events(sizes, world)]
        assert (max(loads) - min(loads)) <= (max(contiguous) - min(contiguous))        # never worse than pulse-balanced ranges
    # fewer heavy events than ranks: nobody is left empty (an empty shard divides by zero in the loss)
    parts = assign_events([5000, 3, 2, 2], 4)
    assert sorted(len(p) for p in parts) == [1, 1, 1, 1]
    assert all(hi > lo for lo, hi in shard_events([100, 1, 1, 1], 4))
    with pytest.raises(ValueError):
        assign_events([5, 4], 3)
    with pytest.raises(ValueError):
        shard_events([5, 4], 3)
