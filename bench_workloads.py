"""The BASELINE configurations that are not the headline line of bench.py (each prints its own JSON line):

  prometheus50  configs[0]: the reference's example data (50 Prometheus events, F = 4, batch 16; examples/04_training/
                01_train_dynedge.py:85,113-142,195,223) -- KNNGraph(Prometheus) -> DynEdge(4) -> energy head + LogCosh ->
                backward -> Adam, one epoch = 4 steps (16 + 16 + 16 + 2 events)
  highmult20k   configs[3]: DynEdge with [min, max, mean, sum] pooling on high-multiplicity events (up to 20 000 pulses)
  tito256       configs[3]: DynEdgeTITO (max aggregation in the GEMM epilogue) on 256 events of <= 256 pulses
  percentile16  configs[3]: the same events as `PercentileClusters` nodes (F = 16), nodes built on the host like the
                reference's dataloader workers (untimed), DynEdge(16) on the device
  microbench    configs[4]: kNN graph build and one DynEdgeConv layer over pulses/event 16 ... 20 000, k = 4 / 8 / 16,
                latent width 128 / 256 / 336 (bit-exactness of edge_index is asserted by tests/test_gpu_knn.py, not here:
                bench.py may not call the oracle outside its CPU-baseline leg)
Run through `python bench.py --workload NAME`.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

# bench.py runs as __main__ and has redirected fd 1 by the time this module is imported: a second copy of it (`import
# bench`) would not know the real stdout and its `_emit` would write the result line to stderr
bench = sys.modules["__main__"] if hasattr(sys.modules.get("__main__"), "_emit") else __import__("bench")

ROOT = os.path.dirname(os.path.abspath(__file__))


def _time_steps(fn, steps, warmup, flush=None):
    """Median CUDA-event time of one call of fn (seconds), L2 flushed between timed calls when `flush` is given."""
    for _ in range(max(warmup, 1)):
        fn()
    torch.cuda.synchronize()
    times = []
    for i in range(steps):
        if flush is not None:
            flush.fill_(float(i))
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(times)), times


class EnergyTrainer:
    """KNNEdges -> DynEdge(F) -> EnergyReconstruction + LogCosh on log10 -> backward -> Adam (configs[0] / [3])."""

    def __init__(self, dev, nb_inputs, backbone=None):
        from graphnet_b200 import ops
        from graphnet_b200.distributed import FlatAdam, FlatGradAllReduce
        from graphnet_b200.models.gnn import DynEdge
        from graphnet_b200.models.graphs.edges import KNNEdges
        from graphnet_b200.tasks import EnergyReconstruction
        torch.manual_seed(0)
        self.backbone = (DynEdge(nb_inputs, global_pooling_schemes=bench.POOLS) if backbone is None else backbone).to(dev)
        self.energy = EnergyReconstruction(128).to(dev)
        self.edges = KNNEdges(8)
        self.reducer = FlatGradAllReduce(list(self.backbone.parameters()) + list(self.energy.parameters()))
        ops.ACCUMULATE_INTO_GRAD = True
        self.opt = FlatAdam(self.reducer, lr=1e-3, eps=1e-3)

    def data(self, db):
        from graphnet_b200 import Data
        return self.edges(Data(x=db["x"], batch=db["batch"], n_pulses=db["n_pulses"]))

    def train_step(self, db):
        pred = self.energy(self.backbone(self.data(db)))
        loss = self.energy.compute_loss(pred, db["energy"])
        loss.backward()
        self.opt.step(zero_grad=True)
        return loss

    @torch.no_grad()
    def infer_step(self, db):
        return self.energy(self.backbone(self.data(db)))


def _line(args, metric, value, sec, config, extra):
    line = {"metric": metric, "value": round(value, 2), "unit": bench.UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(sec * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": bench.DTYPE[args.precision], "data": config.pop("data", "synthetic"), "config": config}
    line.update(extra)
    bench._emit(line)


def prometheus_batches(batch_size=16):
    """The 50 events of the reference's example data base in dataloader order, standardised by the Prometheus detector
    (fixture written by tests/golden/make_prometheus_fixture.py from the reference's own data + detector code)."""
    from graphnet_b200.data import Batch
    from graphnet_b200.models.detector import Prometheus
    from graphnet_b200.models.graphs import KNNGraph
    fx = np.load(os.path.join(ROOT, "tests", "golden", "prometheus_events.npz"))
    names = [str(s) for s in fx["features"]]
    definition = KNNGraph(detector=Prometheus(), input_feature_names=names)
    starts = np.concatenate([[0], np.cumsum(fx["n_pulses"].astype(np.int64))])
    graphs = [definition(fx["raw"][starts[i]:starts[i + 1]], names, truth_dicts=[{"total_energy": float(fx["total_energy"][i])}])
              for i in range(len(fx["n_pulses"]))]
    return [Batch.from_data_list(graphs[i:i + batch_size]) for i in range(0, len(graphs), batch_size)], definition


def run_prometheus50(args, dev):
    from graphnet_b200 import ops
    ops.set_precision(args.precision)
    t0 = time.perf_counter()
    batches, _ = prometheus_batches(16)
    host_s = time.perf_counter() - t0
    dbs = [{"x": b.x.to(dev), "batch": b.batch.to(dev), "n_pulses": b.n_pulses.to(dev), "energy": b.total_energy.float().to(dev)}
           for b in batches]
    tr = EnergyTrainer(dev, 4)

    def epoch():
        for db in dbs:
            tr.train_step(db)

    def epoch_infer():
        for db in dbs:
            tr.infer_step(db)
    sec, _ = _time_steps(epoch, args.steps, args.warmup)
    sec_i, _ = _time_steps(epoch_infer, args.steps, args.warmup)
    n_ev = sum(int(db["n_pulses"].numel()) for db in dbs)
    _line(args, "dynedge_train_events_per_sec", n_ev / sec, sec,
          {"workload": "BASELINE configs[0]: the reference's 50 Prometheus example events (1 872 pulses, F = 4), batch 16 -> 4 steps per "
                       "epoch (16 + 16 + 16 + 2), KNNGraph(Prometheus) -> DynEdge(4) -> energy (LogCosh) fwd + bwd + Adam; one 'step' "
                       "here = one epoch", "precision": f"{args.precision}: {bench.TOLERANCE[args.precision]}",
           "data": "reference example data (tests/golden/prometheus_events.npz)", "events": n_ev, "batches": len(dbs)},
          {"inference": {"value": round(n_ev / sec_i, 2), "unit": bench.UNIT, "ms_per_epoch": round(sec_i * 1e3, 4)},
           "host_graph_definition_ms": round(host_s * 1e3, 2),
           "note": "launch-bound: ~100 launches per 16-event step; parity on this data: tests/test_gpu_config0.py"})


def highmult_batch(seed=11, nev=64, n_max=20000):
    from graphnet_b200.synthetic import event_sizes, make_batch
    rng = np.random.default_rng(seed)
    sizes = event_sizes(nev, rng, sigma=1.5, n_max=n_max)
    sizes[:3] = [20000, 10000, 5000]                       # fixed n in {5k, 10k, 20k} present
    return make_batch(nev, seed=seed, sizes=sizes)


def run_highmult20k(args, dev):
    from graphnet_b200 import ops
    ops.set_precision(args.precision)
    raw = highmult_batch()
    db = {k: torch.from_numpy(np.ascontiguousarray(raw[k])).to(dev) for k in ("x", "batch", "n_pulses", "energy")}
    tr = EnergyTrainer(dev, 7)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    sec, _ = _time_steps(lambda: tr.train_step(db), args.steps, args.warmup, flush)
    sec_i, _ = _time_steps(lambda: tr.infer_step(db), args.steps, args.warmup, flush)
    sec_knn, _ = _time_steps(lambda: tr.data(db), args.steps, args.warmup, flush)
    nev, n = int(db["n_pulses"].numel()), int(db["x"].shape[0])
    _line(args, "dynedge_train_events_per_sec", nev / sec, sec,
          {"workload": "BASELINE configs[3]: DynEdge (pooling min/max/mean/sum) on high-multiplicity events: 64 events, pulses/event "
                       "lognormal(median 100, sigma 1.5) clipped to 20 000 with one 20 000-, one 10 000- and one 5 000-pulse event, "
                       "NodesAsPulses, energy (LogCosh) fwd + bwd + Adam", "precision": f"{args.precision}: {bench.TOLERANCE[args.precision]}",
           "events": nev, "pulses": n, "largest_event": int(raw["n_pulses"].max()), "l2": "256 MiB flush between timed steps"},
          {"pulses_per_s": round(n / sec, 1),
           "inference": {"value": round(nev / sec_i, 2), "unit": bench.UNIT, "ms_per_step": round(sec_i * 1e3, 4), "pulses_per_s": round(n / sec_i, 1)},
           "initial_knn_ms": round(sec_knn * 1e3, 4), "knn_pairs_per_s": round(float((raw["n_pulses"].astype(np.float64) ** 2).sum()) / sec_knn, 1)})


def run_tito256(args, dev):
    """configs[3], the DynEdgeTITO half: 256 events clipped to 256 pulses (the TITO solution caps the pulses per event; its
    per-event TransformerEncoder works on a padded dense [events, longest event, 256] batch)."""
    from graphnet_b200 import ops
    from graphnet_b200.models.gnn import DynEdgeTITO
    from graphnet_b200.synthetic import event_sizes, make_batch
    ops.set_precision(args.precision)
    nev = 256
    sizes = event_sizes(nev, np.random.default_rng(5), sigma=1.0, n_max=256)
    raw = make_batch(nev, seed=5, sizes=sizes)
    db = {k: torch.from_numpy(np.ascontiguousarray(raw[k])).to(dev) for k in ("x", "batch", "n_pulses", "energy")}
    torch.manual_seed(0)
    tr = EnergyTrainer(dev, 7, backbone=DynEdgeTITO(7, global_pooling_schemes=bench.POOLS))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    res = {}
    for name, on in (("max_epilogue", True), ("unfused_max", False)):
        ops.MAX_EPILOGUE = on
        sec, _ = _time_steps(lambda: tr.train_step(db), args.steps, args.warmup, flush)
        sec_i, _ = _time_steps(lambda: tr.infer_step(db), args.steps, args.warmup, flush)
        res[name] = (sec, sec_i)
    ops.MAX_EPILOGUE = True
    sec, sec_i = res["max_epilogue"]
    n = int(db["x"].shape[0])
    _line(args, "dynedge_train_events_per_sec", nev / sec, sec,
          {"workload": "BASELINE configs[3]: DynEdgeTITO (4 DynTrans layers (256, 256), max aggregation, per-event TransformerEncoder, "
                       "pooling min/max/mean/sum) on 256 events of <= 256 pulses, energy (LogCosh) fwd + bwd + Adam; the EdgeConvTito "
                       "MLPs on the kernels (hoisted first Linear, max-aggregating tcgen05 epilogue, arg-routed backward), LayerNorm "
                       "and the transformer as torch modules",
           "precision": f"{args.precision}: {bench.TOLERANCE[args.precision]}", "events": nev, "pulses": n,
           "l2": "256 MiB flush between timed steps"},
          {"pulses_per_s": round(n / sec, 1),
           "inference": {"value": round(nev / sec_i, 2), "unit": bench.UNIT, "ms_per_step": round(sec_i * 1e3, 4)},
           "without_max_epilogue": {"train_ms_per_step": round(res["unfused_max"][0] * 1e3, 4),
                                    "infer_ms_per_step": round(res["unfused_max"][1] * 1e3, 4),
                                    "note": "the second Linear / LeakyReLU of every EdgeConvTito as torch modules on the [E, 256] "
                                            "edge tensor + the SIMT max-aggregation kernel (the route before the fused epilogue)"}})


def run_percentile16(args, dev):
    from graphnet_b200 import ops
    from graphnet_b200.data import Batch
    from graphnet_b200.models.detector import IdentityDetector
    from graphnet_b200.models.graphs import KNNGraph
    from graphnet_b200.models.graphs.nodes import PercentileClusters
    from graphnet_b200.synthetic import FEATURES_ICECUBE86
    ops.set_precision(args.precision)
    raw = highmult_batch()
    definition = KNNGraph(detector=IdentityDetector(), input_feature_names=FEATURES_ICECUBE86, nb_nearest_neighbours=8,
                          node_definition=PercentileClusters(["dom_x", "dom_y", "dom_z"], [10, 50, 90]))
    t0 = time.perf_counter()
    graphs = [definition(raw["x"][raw["ptr"][i]:raw["ptr"][i + 1]], FEATURES_ICECUBE86) for i in range(len(raw["n_pulses"]))]
    host = Batch.from_data_list(graphs)
    host_s = time.perf_counter() - t0
    db = {"x": host.x.to(dev), "batch": host.batch.to(dev), "n_pulses": host.n_pulses.to(dev),
          "energy": torch.from_numpy(raw["energy"]).to(dev)}
    tr = EnergyTrainer(dev, 16)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    sec, _ = _time_steps(lambda: tr.train_step(db), args.steps, args.warmup, flush)
    sec_i, _ = _time_steps(lambda: tr.infer_step(db), args.steps, args.warmup, flush)
    nev, n = int(db["n_pulses"].numel()), int(db["x"].shape[0])
    _line(args, "dynedge_train_events_per_sec", nev / sec, sec,
          {"workload": "BASELINE configs[3]: the high-multiplicity events as PercentileClusters([dom_x, dom_y, dom_z], [10, 50, 90]) nodes "
                       "(F = 3 + 4*3 + 1 = 16, <= 5 160 nodes/event) built per event on the host (untimed, like the reference's dataloader "
                       "workers), DynEdge(16) energy (LogCosh) fwd + bwd + Adam on the device",
           "precision": f"{args.precision}: {bench.TOLERANCE[args.precision]}", "events": nev, "nodes": n,
           "raw_pulses": int(raw["x"].shape[0]), "l2": "256 MiB flush between timed steps"},
          {"nodes_per_s": round(n / sec, 1), "host_node_definition_ms": round(host_s * 1e3, 1),
           "inference": {"value": round(nev / sec_i, 2), "unit": bench.UNIT, "ms_per_step": round(sec_i * 1e3, 4)}})


def run_microbench(args, dev):
    """kNN graph build: total ~160 k pulses split into equal events of n pulses, k in {4, 8, 16}; DynEdgeConv layer (latent
    256 -> hidden C -> 256, k = 8, ~80 k nodes) forward and forward + backward for C in {128, 256, 336}."""
    from graphnet_b200 import ops
    from graphnet_b200.models.components.layers import DynEdgeConv
    ops.set_precision(args.precision)
    rng = np.random.default_rng(3)
    knn_rows = []
    for n_ev in (16, 64, 256, 1024, 4096, 20000):
        nev = max(1, 163840 // n_ev)
        n = nev * n_ev
        x = torch.from_numpy((np.round(rng.normal(size=(n, 3)) * 64) / 64).astype(np.float32)).to(dev)   # ties on purpose
        ptr = torch.arange(0, n + 1, n_ev, dtype=torch.int64, device=dev)
        for k in (4, 8, 16):
            sec, _ = _time_steps(lambda: ops.knn_table(x, [0, 1, 2], ptr, k), max(3, args.steps // 2), 2)
            knn_rows.append({"pulses_per_event": n_ev, "events": nev, "k": k, "ms": round(sec * 1e3, 4),
                             "pairs_per_s": round(float(nev) * n_ev * n_ev / sec, 1), "nodes_per_s": round(n / sec, 1)})
    conv_rows = []
    n, n_ev = 81920, 128
    ptr = torch.arange(0, n + 1, n_ev, dtype=torch.int64, device=dev)
    feat = torch.randn(n, 256, device=dev)
    graph = ops.knn_table(feat, [0, 1, 2], ptr, 8)
    e_real = int(graph.deg.sum().item())
    for width in (128, 256, 336):
        torch.manual_seed(width)
        nn = torch.nn.Sequential(torch.nn.Linear(512, width), torch.nn.ReLU(), torch.nn.Linear(width, 256), torch.nn.ReLU())
        conv = DynEdgeConv(nn, aggr="add", nb_neighbors=8, features_subset=slice(0, 3)).to(dev)
        xg = feat.clone().requires_grad_(True)

        def fwd():
            with torch.no_grad():
                conv.forward_table(feat, graph, ptr, recompute=True)

        def fwd_bwd():
            y, _ = conv.forward_table(xg, graph, ptr, recompute=True)
            y.sum().backward()
        sf, _ = _time_steps(fwd, max(3, args.steps // 2), 2)
        sb, _ = _time_steps(fwd_bwd, max(3, args.steps // 2), 2)
        lit = 2.0 * e_real * (512 * width + width * 256)
        conv_rows.append({"hidden": width, "nodes": n, "edges": e_real, "fwd_ms": round(sf * 1e3, 4), "fwd_bwd_ms": round(sb * 1e3, 4),
                          "fwd_literal_tflops": round(lit / sf / 1e12, 1), "fwd_bwd_literal_tflops": round(3 * lit / sb / 1e12, 1)})
    best = max(knn_rows, key=lambda r: r["nodes_per_s"])
    _line(args, "knn_nodes_per_sec", best["nodes_per_s"], best["ms"] * 1e-3,
          {"workload": "BASELINE configs[4]: kNN graph-build + DynEdgeConv layer microbench sweep (pulses/event 16 ... 20 000, "
                       "k = 4 / 8 / 16, hidden width 128 / 256 / 336); per-operator route (autograd Functions over the C ABI)",
           "precision": f"{args.precision}: {bench.TOLERANCE[args.precision]}"},
          {"unit": "nodes/s", "knn": knn_rows, "dynedgeconv_layer": conv_rows,
           "note": "edge_index bit-exactness over the same grid: tests/test_gpu_knn.py (torch.equal vs both oracles)"})


def run(args, dev, world, rank):
    if rank != 0:
        return
    {"prometheus50": run_prometheus50, "highmult20k": run_highmult20k, "percentile16": run_percentile16,
     "microbench": run_microbench, "tito256": run_tito256}[args.workload](args, dev)
