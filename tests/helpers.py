"""Shared helpers for the parity tests (oracle side = CPU, checker only)."""

from __future__ import annotations

import glob
import os
from types import SimpleNamespace
from typing import Dict, List

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files() -> List[str]:
    # DynEdge cases only (losses.pt holds the task-loss vectors of tests/test_tasks.py, detector_*.pt the standardisation)
    return sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))
                  if os.path.basename(f) != "losses.pt" and not os.path.basename(f).startswith(("detector_", "nodes_", "users_")))


def load_golden(path: str) -> Dict:
    return torch.load(path, map_location="cpu", weights_only=False)


def seeded_state_dict(module: torch.nn.Module, seed: int):
    """Same construction-order independent weights as tests/golden/make_golden.py."""
    out = {}
    for i, (key, val) in enumerate(sorted(module.state_dict().items())):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if key.endswith("weight") and val.dim() == 2:
            out[key] = (torch.rand(val.shape, generator=g) * 2 - 1) / (val.shape[1] ** 0.5)
        elif key.endswith("weight"):
            out[key] = 1.0 + 0.1 * (torch.rand(val.shape, generator=g) * 2 - 1)
        else:
            out[key] = 0.1 * (torch.rand(val.shape, generator=g) * 2 - 1)
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_inf / max(||b||_inf, eps): the per-tensor metric of SURVEY.md section 8d."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def tie_heavy_events(sizes, nb_inputs: int, seed: int):
    """Coarse xyz grid => many exact distance ties and duplicate positions."""
    rng = np.random.default_rng(seed)
    xs = []
    for n in sizes:
        doms = rng.integers(0, 6, size=(max(1, int(np.ceil(0.6 * n))), 3)).astype(np.float32) * 0.25
        pick = rng.integers(0, doms.shape[0], size=n)
        rest = rng.normal(size=(n, nb_inputs - 3)).astype(np.float32)
        xs.append(np.concatenate([doms[pick], rest], axis=1))
    x = torch.from_numpy(np.concatenate(xs, 0))
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    n_pulses = torch.tensor(sizes, dtype=torch.int32)
    return x, batch, n_pulses


def namespace(**kw) -> SimpleNamespace:
    return SimpleNamespace(**kw)
