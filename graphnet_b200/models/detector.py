"""Minimal `Detector`: per-feature standardisation applied before node definition.

Reference: src/graphnet/models/detector/detector.py:63-77 (apply a callable per named column) and
src/graphnet/models/detector/icecube.py:21-48 (IceCube86 constants). Elementwise host-side
preprocessing; only present so `KNNGraph(detector=...)` keeps its constructor contract and so the
benchmark's synthetic pulse maps are standardised like the reference's.
"""

from __future__ import annotations

from typing import Callable, Dict, List

import torch

from graphnet_b200.models.model import Model


class Detector(Model):
    def feature_map(self) -> Dict[str, Callable]:
        raise NotImplementedError

    def forward(self, input_features: torch.Tensor, input_feature_names: List[str]) -> torch.Tensor:
        fmap = self.feature_map()
        out = input_features.clone()
        for idx, name in enumerate(input_feature_names):
            if name in fmap:
                out[:, idx] = fmap[name](input_features[:, idx])
        return out


class IceCube86(Detector):
    """xyz/500, (t-1e4)/3e4, log10(charge), (rde-1.25)/0.25, pmt_area/0.05 (icecube.py:21-48)."""

    def feature_map(self) -> Dict[str, Callable]:
        return {
            "dom_x": lambda v: v / 500.0, "dom_y": lambda v: v / 500.0, "dom_z": lambda v: v / 500.0,
            "dom_time": lambda v: (v - 1.0e04) / 3.0e4, "charge": lambda v: torch.log10(v),
            "rde": lambda v: (v - 1.25) / 0.25, "pmt_area": lambda v: v / 0.05,
        }


class IdentityDetector(Detector):
    def feature_map(self) -> Dict[str, Callable]:
        return {}
