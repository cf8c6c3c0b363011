"""Task heads and losses used by the training benchmark (BASELINE config #3): the O(B) step right
after the hot path. Plain torch ops on the device -- listed as a "next" row in SURVEY.md section 8f.

Reference: `EnergyReconstruction` (src/graphnet/models/task/reconstruction.py:101-112) with
`LogCoshLoss` (src/graphnet/training/loss_functions.py:93-112) on log10(energy) as in
examples/04_training/01_train_dynedge.py:113-124; `DirectionReconstructionWithKappa`
(reconstruction.py:49-70) with `VonMisesFisher3DLoss` (loss_functions.py:281-353, 424-447). Below the reference's
kappa_switch = 100 log C_3(kappa) is evaluated in closed form (kappa / (4 pi sinh kappa)), which is what the reference's
scipy-Bessel `LogCMK` computes for m = 3 (tests/training/test_loss_functions.py:81-83) without the host round trip; from
100 on it follows the reference's shifted approximation (loss_functions.py:307-326).
"""

from __future__ import annotations

import math

import torch
from torch import Tensor


def _eps_like(x: Tensor) -> float:
    return torch.finfo(x.dtype).eps


class EnergyReconstruction(torch.nn.Module):
    def __init__(self, hidden_size: int):
        super().__init__()
        self._affine = torch.nn.Linear(hidden_size, 1)

    def forward(self, x: Tensor) -> Tensor:
        z = self._affine(x)
        return torch.nn.functional.softplus(z, beta=0.05) + _eps_like(z)

    def compute_loss(self, pred: Tensor, energy: Tensor) -> Tensor:
        diff = torch.log10(pred.squeeze(1)) - torch.log10(energy)
        return torch.mean(diff + torch.nn.functional.softplus(-2.0 * diff) - math.log(2.0))


class DirectionReconstructionWithKappa(torch.nn.Module):
    def __init__(self, hidden_size: int):
        super().__init__()
        self._affine = torch.nn.Linear(hidden_size, 3)

    def forward(self, x: Tensor) -> Tensor:
        z = self._affine(x)
        kappa = torch.linalg.vector_norm(z, dim=1) + _eps_like(z)
        return torch.cat([z / kappa.unsqueeze(1), kappa.unsqueeze(1)], dim=1)

    # VonMisesFisherLoss.log_cmk (loss_functions.py:307-326): exact form below kappa_switch = 100, the approximation of
    # arXiv:1812.04616 sec. 8.2 above it, shifted by `offset = approx(100) - exact(100)` for continuity. For m = 3 the
    # approximation is -sqrt(4 + kappa^2) (its log term has the factor m / 2 - 1.5 = 0).
    # The reference evaluates the offset on a float32 `kappa_switch` tensor (loss_functions.py:318-323), whatever the dtype
    # of kappa: -2.78729248046875 (the float64 value would be -2.787291119978647); kept bit for bit.
    KAPPA_SWITCH = 100.0
    LOG_C3_OFFSET = -2.78729248046875

    @classmethod
    def log_c3(cls, kappa: Tensor) -> Tensor:
        # exact: log(kappa / (4 pi sinh kappa)) = log k - log(2 pi) - k - log(1 - exp(-2k)); -expm1 keeps the last term
        # accurate for tiny kappa in fp32
        exact = torch.log(kappa) - math.log(2.0 * math.pi) - kappa - torch.log(-torch.expm1(-2.0 * kappa))
        approx = -torch.sqrt(4.0 + kappa * kappa) - cls.LOG_C3_OFFSET
        return torch.where(kappa < cls.KAPPA_SWITCH, exact, approx)

    def compute_loss(self, pred: Tensor, direction: Tensor) -> Tensor:
        kappa = pred[:, 3]
        p = kappa.unsqueeze(1) * pred[:, :3]
        k = torch.norm(p, dim=1)
        return torch.mean(-self.log_c3(k) - torch.sum(p * direction.reshape(-1, 3), dim=1))


class FusedEnergyDirectionTask(torch.nn.Module):
    """Both heads and both losses of BASELINE config #3 in two CUDA kernels (csrc/task_heads.cu): the same arithmetic as
    `EnergyReconstruction` + log-cosh and `DirectionReconstructionWithKappa` + vMF-3D above, one warp per event, no
    intermediate tensors. `forward(h, energy, direction)` returns `(loss_energy + loss_direction, pred_energy[B, 1],
    pred_direction[B, 4])`. The two plain-torch heads are held as submodules so `state_dict()` keys stay those of the
    separate tasks (`energy._affine.*`, `direction._affine.*`)."""

    def __init__(self, hidden_size: int):
        super().__init__()
        self.energy = EnergyReconstruction(hidden_size)
        self.direction = DirectionReconstructionWithKappa(hidden_size)

    def forward(self, h: Tensor, energy: Tensor, direction: Tensor):
        from graphnet_b200 import ops
        if not h.is_cuda:
            raise RuntimeError("FusedEnergyDirectionTask needs CUDA tensors (no CPU fallback)")
        loss, pe, pd = ops.task_heads_loss(h, self.energy._affine.weight, self.energy._affine.bias,
                                           self.direction._affine.weight, self.direction._affine.bias, energy, direction)
        return loss.sum(), pe.unsqueeze(1), pd
