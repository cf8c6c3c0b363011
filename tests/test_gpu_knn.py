"""Bit-exact parity of the batched kNN kernel against the oracle (through the C ABI)."""

import numpy as np
import pytest
import torch

from helpers import tie_heavy_events
from oracle import c_oracle
from oracle.dynedge_oracle import batch_to_ptr, knn_graph_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 0], ids=["thread_per_query", "auto_split8"])
def knn_variant(request, built_library):
    """Every test runs on the one-thread-per-query kernel and on the default selection (8 lanes per query + merge for
    k = 8 on 3 columns): both must reproduce the oracle's table bit for bit."""
    from graphnet_b200 import _lib
    lib = _lib.load()
    assert lib.gnb_knn_set_variant(request.param) == 0
    yield request.param
    lib.gnb_knn_set_variant(0)


def _kernel_graph(x, cols, ptr, k):
    from graphnet_b200 import ops
    return ops.knn_table(x.cuda(), cols, ptr.cuda(), k)


@pytest.mark.parametrize("k", [1, 4, 8, 16, 23])
def test_knn_bit_exact_tie_heavy(built_library, k):
    sizes = [1, 2, 3, 9, 10, 50, 300, 7, 12, 40, 129, 128, 127, 1]
    x, batch, _ = tie_heavy_events(sizes, 7, seed=k)
    ptr = batch_to_ptr(batch)
    g = _kernel_graph(x, [0, 1, 2], ptr, k)
    ei_ref = knn_graph_ref(x[:, :3], k, ptr=ptr)
    assert torch.equal(g.edge_index().cpu(), ei_ref)
    nbr_c, deg_c = c_oracle.knn_table(x.numpy(), [0, 1, 2], ptr.numpy(), k)
    assert np.array_equal(g.nbr.cpu().numpy(), nbr_c) and np.array_equal(g.deg.cpu().numpy(), deg_c)


def test_knn_duplicate_quirk_and_empty_segments(built_library):
    x = torch.zeros(30, 3)
    x[12:] = torch.rand(18, 3)
    ptr = torch.tensor([0, 12, 12, 13, 30, 30])
    g = _kernel_graph(x, [0, 1, 2], ptr, 8)
    assert g.deg.cpu().tolist()[:13] == [8] * 9 + [9] * 3 + [0]
    assert torch.equal(g.edge_index().cpu(), knn_graph_ref(x, 8, ptr=ptr))


@pytest.mark.parametrize("cols", [[3], [4, 0, 5, 2], list(range(40))])
def test_knn_generic_columns_and_strided_rows(built_library, cols):
    rng = np.random.default_rng(len(cols))
    wide = torch.from_numpy(rng.integers(-2, 3, size=(150, 64)).astype(np.float32) * 0.5)
    x = wide[:, :48]                                        # row pitch 64, width 48
    ptr = torch.tensor([0, 20, 90, 150])
    g = _kernel_graph(wide.cuda()[:, :48], cols, ptr, 5)
    assert torch.equal(g.edge_index().cpu(), knn_graph_ref(x[:, cols], 5, ptr=ptr))


def test_knn_latent_layout_large_events(built_library):
    # columns 0..2 of a 256-wide latent tensor, events up to 5000 nodes (C oracle as checker)
    rng = np.random.default_rng(7)
    sizes = [5000, 3, 1200, 64, 2500]
    n = sum(sizes)
    feat = torch.from_numpy(rng.normal(size=(n, 256)).astype(np.float32))
    feat[:, :3] = torch.round(feat[:, :3] * 8) / 8            # force ties
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]))
    g = _kernel_graph(feat, [0, 1, 2], ptr, 8)
    nbr_c, deg_c = c_oracle.knn_table(feat.numpy(), [0, 1, 2], ptr.numpy(), 8, threads=8)
    assert np.array_equal(g.nbr.cpu().numpy(), nbr_c) and np.array_equal(g.deg.cpu().numpy(), deg_c)


def test_knn_full_size_properties(built_library):
    # BASELINE config #2 size (B=1024): properties that do not need the oracle at full size
    from graphnet_b200 import ops
    from graphnet_b200.synthetic import make_batch
    raw = make_batch(1024, seed=20240607)
    x = torch.from_numpy(raw["x"]).cuda()
    ptr = torch.from_numpy(raw["ptr"]).cuda()
    batch = torch.from_numpy(raw["batch"]).cuda()
    g = ops.knn_table(x, [0, 1, 2], ptr, 8)
    ei = g.edge_index()
    assert torch.equal(batch[ei[0]], batch[ei[1]])                      # never crosses events
    assert not bool((ei[0] == ei[1]).any())                            # no self loops
    n_b = (ptr[1:] - ptr[:-1])[batch]
    deg = g.deg.long()
    assert bool(((deg == torch.clamp(n_b - 1, max=8)) | (deg == 9)).all())
    d = ((x[ei[0], :3] - x[ei[1], :3]) ** 2).sum(1)
    same_q = ei[1][1:] == ei[1][:-1]
    assert bool((d[1:][same_q] >= d[:-1][same_q] - 1e-6).all())          # ascending distance per query
    # idempotence / determinism
    g2 = ops.knn_table(x, [0, 1, 2], ptr, 8)
    assert torch.equal(g.nbr, g2.nbr)
    # a sample of events against the C oracle
    sel = [0, 1, 2, 3, 500, 1023]
    for b in sel:
        lo, hi = int(raw["ptr"][b]), int(raw["ptr"][b + 1])
        nbr_c, deg_c = c_oracle.knn_table(raw["x"][lo:hi], [0, 1, 2], np.array([0, hi - lo]), 8)
        got = g.nbr[lo:hi].cpu().numpy()
        assert np.array_equal(np.where(got >= 0, got - lo, -1), nbr_c)


@pytest.mark.gpu
def test_device_graph_definition_from_raw_pulses(built_library):
    """raw pulse batch on the GPU -> standardised nodes, batch / ptr / n_pulses, kNN table in three
    launches. x is checked against the golden output of the reference's own detector/icecube.py (CPU fp32): affine
    columns bit-exact (IEEE subtract / divide), log10(charge) within 2 ulp (CUDA log10f vs the CPU libm the reference's
    dataloader workers use; tolerance rel 2.4e-7 + the value's own ulp); batch / ptr equal the collate restatement and the
    host path (per-event KNNGraph.forward + Batch.from_data_list); the graph equals the oracle kNN on the device's x."""
    import os
    from helpers import GOLDEN_DIR
    from graphnet_b200.data import Batch
    from graphnet_b200.models.detector import IceCube86
    from graphnet_b200.models.graphs import DeviceKNNGraph, KNNGraph
    from oracle.dynedge_oracle import collate_ref
    gold = torch.load(os.path.join(GOLDEN_DIR, "detector_icecube86.pt"))
    raw, want, names = gold["raw"], gold["standardized"], gold["features"]
    sizes = torch.tensor([1, 2, 9, 10, 300, 64, 700, 0, 1500, 1510], dtype=torch.int32)   # sums to 4096; one empty event
    assert int(sizes.sum()) == raw.shape[0]
    definition = KNNGraph(detector=IceCube86(), input_feature_names=names, nb_nearest_neighbours=8, columns=[0, 1, 2])
    graph = DeviceKNNGraph(definition)(raw.cuda(), sizes.cuda())
    x = graph.x.cpu()
    for c, name in enumerate(names):
        if name == "charge":
            assert torch.allclose(x[:, c], want[:, c], rtol=2.4e-7, atol=1e-9), name
        else:
            assert torch.equal(x[:, c], want[:, c]), name
        assert torch.equal(graph[name].cpu(), x[:, c])                # per-feature attributes, graph_definition.py:243-247
    batch_ref, ptr_ref = collate_ref(sizes)
    assert torch.equal(graph.batch.cpu(), batch_ref) and torch.equal(graph.ptr.cpu(), ptr_ref)
    assert torch.equal(graph.n_pulses.cpu(), sizes) and graph.n_pulses.dtype == torch.int32
    assert torch.equal(graph.edge_index.cpu(), knn_graph_ref(x[:, :3], 8, ptr=ptr_ref))
    # host path of the reference's dataloader: one Data per (non-empty) event, collated
    off, parts = 0, []
    for s_ in sizes.tolist():
        if s_ > 0:
            parts.append(definition(raw[off:off + s_].numpy(), names))
        off += s_
    host = Batch.from_data_list(parts)
    keep = sizes > 0
    assert torch.equal(host.n_pulses, sizes[keep])
    for c, name in enumerate(names):
        if name != "charge":
            assert torch.equal(host.x[:, c], x[:, c])
    with pytest.raises(RuntimeError):                                   # no CPU fallback
        DeviceKNNGraph(definition)(raw, sizes)


@pytest.mark.gpu
def test_device_percentile_clusters_equal_the_host_node_definition(built_library):
    """SURVEY 8f rank 3, second half: `PercentileClusters` nodes built for a whole raw batch on the device
    (models/graphs/device.py::percentile_clusters: segmented lexsort, float64 percentile arithmetic of numpy) against the
    host node definition, which is pinned bit for bit on the reference's own cluster_summarize_with_percentiles
    (tests/test_oracle_golden.py). Nodes bit-exact (log10(count): 1 ulp of fp32), node ptr, raw n_pulses, kNN graph."""
    from graphnet_b200.data import Batch
    from graphnet_b200.models.detector import IceCube86
    from graphnet_b200.models.graphs import DeviceKNNGraph, KNNGraph
    from graphnet_b200.models.graphs.nodes import PercentileClusters
    from graphnet_b200.synthetic import FEATURES_ICECUBE86
    rng = np.random.default_rng(7)
    sizes = [300, 1, 40, 7, 2500, 64]
    raws = []
    for n_ in sizes:
        doms = np.round(rng.uniform(-500, 500, size=(max(1, n_ // 3), 3)), 0).astype(np.float32)
        rest = np.stack([rng.normal(1e4, 1.5e3, n_), rng.lognormal(0.0, 0.7, n_), rng.choice([1.0, 1.35], n_),
                         np.full(n_, 0.0444)], axis=1).astype(np.float32)
        raws.append(np.concatenate([doms[rng.integers(0, len(doms), size=n_)], rest], axis=1))
    definition = KNNGraph(detector=IceCube86(), input_feature_names=FEATURES_ICECUBE86, nb_nearest_neighbours=8,
                          node_definition=PercentileClusters(["dom_x", "dom_y", "dom_z"], [10, 50, 90]))
    host = Batch.from_data_list([definition(r, FEATURES_ICECUBE86) for r in raws])
    raw = torch.from_numpy(np.concatenate(raws))
    n_pulses = torch.tensor(sizes, dtype=torch.int32)
    graph = DeviceKNNGraph(definition)(raw.cuda(), n_pulses.cuda())
    x = graph.x.cpu()
    assert x.shape == host.x.shape and x.shape[1] == 16
    # the charge percentiles go through log10f on the device (2 ulp of values up to ~1, see the NodesAsPulses test: 2.4e-7
    # absolute, also where the interpolated percentile itself is near zero); everything else is exact
    names = definition.output_feature_names
    for c, name in enumerate(names):
        if name.startswith("charge") or name == "counts":
            assert torch.allclose(x[:, c], host.x[:, c], rtol=3e-7, atol=5e-7), name
        else:
            assert torch.equal(x[:, c], host.x[:, c]), name
    assert torch.equal(graph.ptr.cpu(), host.ptr) and torch.equal(graph.batch.cpu(), host.batch)
    assert torch.equal(graph.n_pulses.cpu(), n_pulses)                          # the RAW pulse counts (graph_definition.py:213)
    assert torch.equal(graph.edge_index.cpu(), knn_graph_ref(x[:, :3], 8, ptr=host.ptr))
