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
    below = g["kappa"] < 100.0                 # from kappa_switch = 100 on the reference uses its approximation (next test)
    g = {key: v[below] for key, v in g.items()}
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


def test_log_c3_follows_the_reference_kappa_switch():
    """loss_functions.py:307-326: above kappa_switch = 100 the reference uses the shifted approximation -- values, gradients
    and the vMF loss itself on predictions with |z| up to 1000 (golden vectors from the reference's own log_cmk)."""
    g = GOLD["log_cmk_switch"]
    k = g["kappa"].clone().requires_grad_(True)
    val = DirectionReconstructionWithKappa.log_c3(k)
    (grad,) = torch.autograd.grad(val.sum(), k)
    assert torch.allclose(val, g["value"], rtol=1e-10, atol=1e-9)
    assert torch.allclose(grad, g["grad"], rtol=1e-7, atol=1e-9)
    head = _identity_head(DirectionReconstructionWithKappa, 3)
    z = g["z"].clone().requires_grad_(True)
    pred = head(z)
    kap = pred[:, 3]
    p = kap.unsqueeze(1) * pred[:, :3]
    el = -DirectionReconstructionWithKappa.log_c3(torch.norm(p, dim=1)) - torch.sum(p * g["target"], dim=1)
    (gz,) = torch.autograd.grad(el.mean(), z)
    assert torch.allclose(el, g["elements"], rtol=1e-10, atol=1e-9)
    assert torch.allclose(gz, g["grad_z"], rtol=1e-7, atol=1e-9)
    # fp32: -expm1 keeps log(1 - exp(-2k)) finite and accurate for tiny kappa
    tiny = torch.tensor([1e-6, 1e-5], dtype=torch.float32)
    ref = (torch.log(tiny.double()) - math.log(2 * math.pi) - tiny.double() - torch.log(-torch.expm1(-2 * tiny.double())))
    assert torch.allclose(DirectionReconstructionWithKappa.log_c3(tiny).double(), ref, rtol=1e-5)
