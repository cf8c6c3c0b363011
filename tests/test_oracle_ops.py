"""Known-answer and cross-formulation tests of the oracle itself (CPU only)."""

import numpy as np
import pytest
import torch

from helpers import tie_heavy_events
from oracle import c_oracle
from oracle.dynedge_oracle import (batch_to_ptr, edgeconv_ref, global_variables_ref, homophily_ref, knn_graph_ref,
                                   segment_pool_ref)


def test_knn_hand_computed():
    # 5 collinear points, k=2: ties resolved towards the lower index
    x = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0], [3.0, 0, 0], [10.0, 0, 0]])
    ei = knn_graph_ref(x, 2)
    expect = {0: [1, 2], 1: [0, 2], 2: [1, 3], 3: [2, 1], 4: [3, 2]}
    for q, nbrs in expect.items():
        assert ei[0][ei[1] == q].tolist() == nbrs
    assert ei[1].tolist() == sorted(ei[1].tolist())          # grouped by target


def test_knn_small_events_and_duplicate_quirk():
    # n < k+1 -> n-1 neighbours; > k exact duplicates at lower index -> k+1 edges (torch_cluster quirk)
    x = torch.zeros(12, 3)
    ei = knn_graph_ref(x, 8)
    deg = torch.bincount(ei[1], minlength=12)
    assert deg.tolist() == [8] * 9 + [9] * 3
    x2 = torch.rand(4, 3)
    assert torch.bincount(knn_graph_ref(x2, 8)[1], minlength=4).tolist() == [3] * 4
    assert knn_graph_ref(torch.rand(1, 3), 8).shape == (2, 0)


@pytest.mark.parametrize("k", [1, 4, 8, 16])
def test_knn_sort_formulation_equals_c_insertion_formulation(k):
    sizes = [1, 2, 3, 9, 10, 50, 300, 7, 12, 40]
    x, batch, _ = tie_heavy_events(sizes, 5, seed=k)
    ptr = batch_to_ptr(batch)
    ei_t = knn_graph_ref(x[:, [0, 1, 2]], k, ptr=ptr)
    ei_c = c_oracle.knn_edge_index(x.numpy(), [0, 1, 2], ptr.numpy(), k, threads=2)
    assert np.array_equal(ei_t.numpy(), ei_c)
    # every edge stays inside its event, no self loops
    assert torch.equal(batch[ei_t[0]], batch[ei_t[1]]) and not bool((ei_t[0] == ei_t[1]).any())


def test_knn_generic_columns():
    x, batch, _ = tie_heavy_events([20, 33], 6, seed=3)
    ptr = batch_to_ptr(batch)
    cols = [4, 0, 5, 2]
    ei_t = knn_graph_ref(x[:, cols], 3, ptr=ptr)
    ei_c = c_oracle.knn_edge_index(x.numpy(), cols, ptr.numpy(), 3)
    assert np.array_equal(ei_t.numpy(), ei_c)


def test_segment_pool_against_c_and_hand():
    x = torch.tensor([[1.0, -2.0], [3.0, -2.0], [2.0, 5.0], [7.0, 7.0]])
    ptr = torch.tensor([0, 3, 3, 4])
    assert segment_pool_ref(x, ptr, "sum").tolist() == [[6.0, 1.0], [0.0, 0.0], [7.0, 7.0]]
    assert segment_pool_ref(x, ptr, "mean").tolist() == [[2.0, float(np.float32(1.0) / np.float32(3.0))], [0.0, 0.0], [7.0, 7.0]]
    assert segment_pool_ref(x, ptr, "max").tolist() == [[3.0, 5.0], [0.0, 0.0], [7.0, 7.0]]
    assert segment_pool_ref(x, ptr, "min").tolist() == [[1.0, -2.0], [0.0, 0.0], [7.0, 7.0]]
    rng = np.random.default_rng(0)
    xr = torch.from_numpy(rng.integers(-3, 4, size=(60, 5)).astype(np.float32))
    ptr = torch.tensor([0, 10, 10, 35, 60])
    for scheme in ("min", "max", "sum", "mean"):
        out_c, _ = c_oracle.segment_pool(xr.numpy(), ptr.numpy(), scheme)
        np.testing.assert_allclose(segment_pool_ref(xr, ptr, scheme).numpy(), out_c, rtol=1e-6, atol=1e-6)


def test_segment_extreme_gradient_goes_to_first_arg():
    x = torch.tensor([[1.0], [5.0], [5.0], [2.0]], requires_grad=True)
    segment_pool_ref(x, torch.tensor([0, 4]), "max").sum().backward()
    assert x.grad.flatten().tolist() == [0.0, 1.0, 0.0, 0.0]


def test_homophily_and_global_variables():
    x, batch, n_pulses = tie_heavy_events([6, 15, 2], 7, seed=5)
    ptr = batch_to_ptr(batch)
    ei = knn_graph_ref(x[:, :3], 4, ptr=ptr)
    g = global_variables_ref(x, ei, batch, n_pulses, ptr)
    assert g.shape == (3, 12)
    for c in range(4):
        h_c = c_oracle.homophily(x.numpy(), c, ei.numpy(), batch.numpy(), 3)
        np.testing.assert_array_equal(g[:, 7 + c].numpy(), h_c)
        np.testing.assert_array_equal(homophily_ref(ei, x[:, c], batch, 3).numpy(), h_c)
    np.testing.assert_allclose(g[:, 11].numpy(), np.log10(np.array([6, 15, 2], dtype=np.float32)), rtol=1e-6)
    np.testing.assert_allclose(g[0, :7].numpy(), x[:6].mean(0).numpy(), rtol=1e-5, atol=1e-6)


def test_edgeconv_hand_computed_and_empty_neighbourhood():
    x = torch.tensor([[1.0], [2.0], [4.0]])
    ei = torch.tensor([[1, 2, 0], [0, 0, 1]])                       # node 2 has no in-edges
    nn = torch.nn.Identity()
    out = edgeconv_ref(x, ei, nn, "add")                            # rows: [x_i, x_j - x_i]
    assert out.tolist() == [[2.0, 1.0 + 3.0], [2.0, -1.0], [0.0, 0.0]]
    assert edgeconv_ref(x, ei, nn, "mean").tolist() == [[1.0, 2.0], [2.0, -1.0], [0.0, 0.0]]
    assert edgeconv_ref(x, ei, nn, "max").tolist() == [[1.0, 3.0], [2.0, -1.0], [0.0, 0.0]]
