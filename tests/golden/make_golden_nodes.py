"""Golden vectors for the `PercentileClusters` node definition, produced by the REFERENCE's own, unmodified
`cluster_summarize_with_percentiles` / `identify_indices` (src/graphnet/models/graphs/utils.py:100-172) -- the functions
`PercentileClusters._construct_nodes` calls (nodes/nodes.py:196-217). Only `graphnet.constants.DATA_DIR` is shimmed
(numpy, pandas, scipy and scikit-learn are installed here).

Run (only in the build container, where /root/reference exists):
    python tests/golden/make_golden_nodes.py
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

FEATURES = ["dom_x", "dom_y", "dom_z", "dom_time", "charge", "rde", "pmt_area"]
CLUSTER_ON = ["dom_x", "dom_y", "dom_z"]
PERCENTILES = [10, 50, 90]


def load_reference_utils():
    mg.install_shims()
    mg._mod("graphnet.constants", DATA_DIR="/nonexistent")
    m = mg._mod("graphnet.models.graphs")
    m.__path__ = [os.path.join(mg.REF_SRC, "graphnet", "models", "graphs")]
    return importlib.import_module("graphnet.models.graphs.utils")


def make_event(rng, n, n_doms):
    doms = np.round(rng.uniform(-1.0, 1.0, size=(n_doms, 3)), 2).astype(np.float32)
    doms[0] = 0.0
    pick = rng.integers(0, n_doms, size=n)
    rest = rng.normal(size=(n, 4)).astype(np.float32)
    return np.concatenate([doms[pick], rest], axis=1)


def main() -> None:
    utils = load_reference_utils()
    cluster_idx, summ_idx, _ = utils.identify_indices(FEATURES, CLUSTER_ON)
    rng = np.random.default_rng(7)
    cases = []
    for n, n_doms in [(1, 1), (2, 1), (5, 5), (40, 6), (300, 60), (1000, 35)]:
        x = make_event(rng, n, n_doms)
        for add_counts in (True, False):
            out = utils.cluster_summarize_with_percentiles(x=x.copy(), summarization_indices=summ_idx, cluster_indices=cluster_idx,
                                                           percentiles=PERCENTILES, add_counts=add_counts)
            cases.append({"x": torch.from_numpy(x), "add_counts": add_counts, "nodes": torch.tensor(out)})
    torch.save({"features": FEATURES, "cluster_on": CLUSTER_ON, "percentiles": PERCENTILES, "cases": cases},
               os.path.join(HERE, "nodes_percentile_clusters.pt"))
    print("wrote nodes_percentile_clusters.pt:", [tuple(c["nodes"].shape) for c in cases])


if __name__ == "__main__":
    main()
