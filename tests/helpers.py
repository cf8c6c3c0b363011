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


# Largest |pre-activation| (relative to the largest one of the tensor) at which the kernel's read-out ReLU decision may
# differ from the oracle's: the forward error of the mode with a safety factor.
FLIP_TOL = {"fp32": 2e-5, "tf32x3": 2e-4, "tf32": 5e-3, "bf16x3": 2e-4, "mixed16": 2e-4, "f16": 5e-3, "bf16": 3e-2}


def oracle_on_kernel_decisions(ref, data, forced_graphs, y_kernel, mode: str):
    """Oracle forward teacher-forced with the kernel's latent graphs AND the kernel's ReLU decisions of the LAST activation
    (y_kernel > 0). The gradient of a ReLU network jumps where a pre-activation crosses zero; at the read-out ([B, 128]
    units) one unit within rounding of zero moves a bias-gradient entry by 1 / B, so gradients can only be compared on the
    same side of every such kink. Asserts that the two sides disagree ONLY within the forward error of `mode` (FLIP_TOL,
    relative to the largest pre-activation). Returns (y_ref, intermediates, number of forced decisions)."""
    mask = (y_kernel.detach().cpu() > 0)
    y_ref, inter = ref(data, forced_graphs=forced_graphs, return_intermediates=True, forced_output_mask=mask)
    z = inter["final_pre"]
    forced = 0
    if z is not None:
        differ = mask != (z.detach() > 0)
        forced = int(differ.sum())
        if forced:
            worst = float(z.detach()[differ].abs().max() / z.detach().abs().max())
            assert worst <= FLIP_TOL[mode], f"read-out ReLU decision differs at |z| = {worst:.2e} of max (mode {mode})"
    return y_ref, inter, forced
