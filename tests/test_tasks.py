"""Task heads + losses against golden vectors produced by the reference's own
`training/loss_functions.py` (tests/golden/make_golden_losses.py) and against the known-answer formulas of the
reference's tests (tests/training/test_loss_functions.py:40-63 log-cosh, :66-95 exact vMF m = 3)."""
import math
import os

import numpy as np
import pytest
import torch

from graphnet_b200.tasks import DirectionReconstructionWithKappa, EnergyReconstruction

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "losses.pt"))


def _identity_head(cls, n):
    head = cls(n).double()
    with torch.no_grad():
        head._affine.weight.copy_(torch.eye(head._affine.out_features, n, dtype=torch.float64))
        head._affine.bias.zero_()
    return head


def test_logcosh_matches_reference_loss_function():
    g = GOLD["logcosh"]
    head = EnergyReconstruction(1).double()
    pl = g["pred_log10"].clone().requires_grad_(True)
    loss = head.compute_loss(10 ** pl, 10 ** g["true_log10"].squeeze(1))
    (grad,) = torch.autograd.grad(loss, pl)
    assert torch.allclose(loss, g["loss"], rtol=1e-12, atol=1e-14)
    assert torch.allclose(grad, g["grad_pred_log10"], rtol=1e-9, atol=1e-12)


def test_logcosh_known_answers_and_stability():
    """Reference test_log_cosh: finite for |x| = 100, equal to log(cosh(x)) wherever that is finite."""
    head = EnergyReconstruction(1)
    x = torch.tensor([-100.0, -10.0, -1.0, 0.0, 1.0, 10.0, 100.0], dtype=torch.float64)
    el = torch.stack([head.compute_loss((10 ** v).reshape(1, 1), torch.ones(1, dtype=torch.float64)) for v in x])
    ref = torch.log(torch.cosh(x))
    assert torch.all(torch.isfinite(el))
    ok = torch.isfinite(ref)
    assert torch.allclose(el[ok], ref[ok], rtol=1e-10, atol=1e-12)


def test_energy_head_transform():
    """reconstruction.py:109-112: softplus(x, beta=0.05) + eps."""
    head = _identity_head(EnergyReconstruction, 1)
    x = torch.tensor([[-300.0], [-1.0], [0.0], [2.0], [500.0]], dtype=torch.float64)
    out = head(x)
    assert torch.all(out > 0)
    assert torch.allclose(out, torch.nn.functional.softplus(x, beta=0.05) + torch.finfo(torch.float64).eps)


def test_vmf3d_matches_reference_loss_function():
    g = GOLD["vmf3d"]
    head = _identity_head(DirectionReconstructionWithKappa, 3)
    z = g["z"].clone().requires_grad_(True)
    pred = head(z)
    assert torch.allclose(pred[:, :3].norm(dim=1), torch.ones(z.shape[0], dtype=torch.float64), atol=1e-12)
    assert torch.allclose(pred[:, 3], z.norm(dim=1) + torch.finfo(torch.float64).eps)
    loss = head.compute_loss(pred, g["target"])
    (grad,) = torch.autograd.grad(loss, z)
    # the reference evaluates log C_3 through scipy Bessel functions; the closed form agrees to ~1e-10 relative
    assert torch.allclose(loss, g["loss"], rtol=1e-9, atol=1e-10)
    assert torch.allclose(grad, g["grad_z"], rtol=1e-6, atol=1e-9)


def test_log_c3_known_answers():
    """Reference test_von_mises_fisher_exact_m3: log k - k - log(2 pi (1 - exp(-2k))), values and gradients."""
    g = GOLD["log_c3"]
    k = g["kappa"].clone().requires_grad_(True)
    val = DirectionReconstructionWithKappa.log_c3(k)
    (grad,) = torch.autograd.grad(val.sum(), k)
    k2 = g["kappa"].clone().requires_grad_(True)
    ref = torch.log(k2) - k2 - torch.log(2 * np.pi * (1 - torch.exp(-2 * k2)))
    (gref,) = torch.autograd.grad(ref.sum(), k2)
    assert torch.allclose(val, ref) and torch.allclose(grad, gref)
    # and the reference's own scipy-Bessel evaluation
    assert torch.allclose(val, g["value"], rtol=1e-8, atol=1e-10)
    assert torch.allclose(grad, g["grad"], rtol=1e-6, atol=1e-8)
    assert math.isfinite(float(DirectionReconstructionWithKappa.log_c3(torch.tensor(1e-4))))
