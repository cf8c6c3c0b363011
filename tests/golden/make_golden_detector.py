"""Golden vectors for the detector standardisation, produced by the REFERENCE's own, unmodified
`src/graphnet/models/detector/icecube.py` (IceCube86.feature_map :21-48) driven through the reference's own
`Detector._standardize` (`detector.py:63-77`), loaded with the package stand-ins of make_golden.py (only
`graphnet.models.Model`, `graphnet.utilities.decorators.final`, `graphnet.constants` and torch_geometric's `Data`
are shimmed). Run on the CPU in fp32 -- the situation inside the reference's dataloader workers.

Run (only in the build container, where /root/reference exists):
    python tests/golden/make_golden_detector.py
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

FEATURES = ["dom_x", "dom_y", "dom_z", "dom_time", "charge", "rde", "pmt_area"]   # FEATURES.ICECUBE86, data/constants.py:7-15


def load_reference_icecube86():
    mg.install_shims()
    mg._mod("graphnet.utilities")
    mg._mod("graphnet.utilities.decorators", final=lambda f: f)
    mg._mod("graphnet.constants", ICECUBE_GEOMETRY_TABLE_DIR="/nonexistent")
    m = mg._mod("graphnet.models.detector")
    m.__path__ = [os.path.join(mg.REF_SRC, "graphnet", "models", "detector")]
    return importlib.import_module("graphnet.models.detector.icecube").IceCube86


def main() -> None:
    cls = load_reference_icecube86()
    det = cls.__new__(cls)                      # the stand-in Model base needs no constructor arguments
    torch.nn.Module.__init__(det)
    rng = np.random.default_rng(20240607)
    n = 4096
    raw = np.stack([rng.uniform(-600, 600, n), rng.uniform(-600, 600, n), rng.uniform(-520, 530, n),
                    rng.normal(1.0e4, 1.5e3, n), rng.lognormal(0.0, 0.7, n),
                    rng.choice([1.0, 1.35], n), rng.choice([0.0444, 0.0222], n)], axis=1).astype(np.float32)
    raw[:8, :3] = 0.0                           # exact zeros / signed zero survive the affine map
    raw[3, 0] = -0.0
    x = torch.from_numpy(raw.copy())
    out = det._standardize(x.clone(), FEATURES)        # the reference writes in place into its argument
    torch.save({"features": FEATURES, "raw": torch.from_numpy(raw), "standardized": out}, os.path.join(HERE, "detector_icecube86.pt"))
    print("wrote detector_icecube86.pt", out.shape, out.dtype, float(out.abs().max()))


if __name__ == "__main__":
    main()
