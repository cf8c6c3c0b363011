"""Golden vectors for the task losses, produced by the REFERENCE's own, unmodified
`src/graphnet/training/loss_functions.py` (LogCoshLoss :93-112, VonMisesFisher3DLoss :424-447 with the scipy-Bessel
`LogCMK` :211-279) loaded through the package stand-ins of make_golden.py (only `graphnet.models.model.Model` and
`graphnet.utilities.decorators.final` are shimmed; scipy is installed here).

Run (only in the build container, where /root/reference exists):
    python tests/golden/make_golden_losses.py
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


def load_reference_losses():
    mg.install_shims()
    mg._mod("graphnet.models.model", Model=mg._Model)
    mg._mod("graphnet.utilities")
    mg._mod("graphnet.utilities.decorators", final=lambda f: f)
    m = mg._mod("graphnet.training")
    m.__path__ = [os.path.join(mg.REF_SRC, "graphnet", "training")]
    return importlib.import_module("graphnet.training.loss_functions")


def main() -> None:
    lf = load_reference_losses()
    g = torch.Generator().manual_seed(11)
    # --- LogCosh on log10(E): predictions in the positive energy domain, as EnergyReconstruction emits them
    n = 64
    pred_e = 10 ** (torch.rand(n, 1, generator=g, dtype=torch.float64) * 5 - 0.5)
    true_e = 10 ** (torch.rand(n, 1, generator=g, dtype=torch.float64) * 4)
    pl = torch.log10(pred_e).requires_grad_(True)
    logcosh = lf.LogCoshLoss()
    el = logcosh(pl, torch.log10(true_e), return_elements=True)
    loss = el.mean()
    (gl,) = torch.autograd.grad(loss, pl)
    # --- vMF 3D: prediction = [unit vector, kappa], target = unit vector
    z = torch.randn(n, 3, generator=g, dtype=torch.float64) * torch.logspace(-2, 1.5, n, dtype=torch.float64).unsqueeze(1)
    z.requires_grad_(True)
    kappa = torch.linalg.vector_norm(z, dim=1) + torch.finfo(torch.float64).eps
    pred_d = torch.cat([z / kappa.unsqueeze(1), kappa.unsqueeze(1)], dim=1)
    t = torch.randn(n, 3, generator=g, dtype=torch.float64)
    t = t / t.norm(dim=1, keepdim=True)
    vmf = lf.VonMisesFisher3DLoss()
    ev = vmf(pred_d, t, return_elements=True)
    lv = ev.mean()
    (gz,) = torch.autograd.grad(lv, z)
    # --- log C_3(kappa) through the reference's exact (scipy Bessel) routine, values + gradients
    k = torch.tensor([1e-4, 1e-3, 1e-2, 0.1, 1.0, 3.0, 10.0, 30.0, 100.0], dtype=torch.float64, requires_grad=True)
    c3 = lf.VonMisesFisherLoss.log_cmk_exact(3, k)
    (gk,) = torch.autograd.grad(c3.sum(), k)
    # --- the full log_cmk (exact below kappa_switch = 100, shifted approximation above), through the vMF loss itself
    kb = torch.tensor([50.0, 99.0, 99.999, 100.0, 100.001, 150.0, 200.0, 1000.0], dtype=torch.float64, requires_grad=True)
    cb = lf.VonMisesFisherLoss.log_cmk(3, kb)
    (gkb,) = torch.autograd.grad(cb.sum(), kb)
    zb = torch.stack([kb.detach(), torch.zeros_like(kb.detach()), torch.zeros_like(kb.detach())], dim=1).requires_grad_(True)
    kap = torch.linalg.vector_norm(zb, dim=1) + torch.finfo(torch.float64).eps
    tb = torch.tensor([[0.6, 0.0, 0.8]], dtype=torch.float64).expand(len(kb), 3)
    evb = vmf(torch.cat([zb / kap.unsqueeze(1), kap.unsqueeze(1)], dim=1), tb, return_elements=True)
    (gzb,) = torch.autograd.grad(evb.mean(), zb)
    out = {"log_cmk_switch": {"kappa": kb.detach(), "value": cb.detach(), "grad": gkb, "z": zb.detach(), "target": tb.clone(),
                              "elements": evb.detach(), "grad_z": gzb},
           "logcosh": {"pred_log10": pl.detach(), "true_log10": torch.log10(true_e), "elements": el.detach(),
                       "loss": loss.detach(), "grad_pred_log10": gl},
           "vmf3d": {"z": z.detach(), "target": t, "elements": ev.detach(), "loss": lv.detach(), "grad_z": gz},
           "log_c3": {"kappa": k.detach(), "value": c3.detach(), "grad": gk}}
    torch.save(out, os.path.join(HERE, "losses.pt"))
    print("wrote losses.pt:", {k_: float(v["loss"]) for k_, v in out.items() if "loss" in v})


if __name__ == "__main__":
    main()
