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
    """bench.host_batches(world > 1): the ranks' shards are the contiguous, pulse-balanced ranges of ONE global batch --
    concatenated they reproduce it exactly (events, pulses, labels), with `batch` renumbered from 0 on every rank."""
    import os
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    world, per_gpu = 3, 16
    full = bench.host_batches(per_gpu * world, 1, seed0=5)[0]
    shards = [bench.host_batches(per_gpu, 1, seed0=5, rank=r, world=world)[0] for r in range(world)]
    assert sum(int(s["n_pulses"].numel()) for s in shards) == per_gpu * world
    assert torch.equal(torch.cat([s["x"] for s in shards]), full["x"])
    assert torch.equal(torch.cat([s["n_pulses"] for s in shards]), full["n_pulses"])
    assert torch.equal(torch.cat([s["energy"] for s in shards]), full["energy"])
    assert torch.equal(torch.cat([s["direction"] for s in shards]), full["direction"])
    pulses = [int(s["x"].shape[0]) for s in shards]
    assert max(pulses) - min(pulses) <= int(full["n_pulses"].max())          # balanced to within one event
    for s in shards:
        b = s["batch"]
        assert int(b.min()) == 0 and int(b.max()) == s["n_pulses"].numel() - 1
        assert torch.equal(torch.bincount(b), s["n_pulses"].long())
