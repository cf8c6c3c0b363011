"""Minimal `Detector`: per-feature standardisation applied before node definition.

Reference: src/graphnet/models/detector/detector.py:63-77 (apply a callable per named column; an unknown
column raises `KeyError`) and src/graphnet/models/detector/icecube.py:21-48 (IceCube86 constants).
`forward` is the host-side form used per event inside dataloader workers; `standardisation_table` exposes
the same map as (kind, subtract, divide) triples so that `models/graphs/device.py` can standardise a whole
raw pulse batch on the GPU in one launch (`gnb_standardize`).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch

from graphnet_b200.models.model import Model

# kinds understood by csrc/graph_ops.cu::standardize_kernel
STD_IDENTITY, STD_AFFINE, STD_LOG10 = 0, 1, 2


class Detector(Model):
    def feature_map(self) -> Dict[str, Callable]:
        raise NotImplementedError

    def forward(self, input_features: torch.Tensor, input_feature_names: List[str]) -> torch.Tensor:
        fmap = self.feature_map()
        out = input_features.clone()
        for idx, name in enumerate(input_feature_names):
            # reference: a missing standardisation function is an error (detector.py:70-76), not a pass-through
            out[:, idx] = fmap[name](input_features[:, idx])
        return out

    def affine_map(self) -> Dict[str, Tuple[int, float, float]]:
        """name -> (kind, subtract, divide): x, (x - subtract) / divide or log10(x). Detectors whose map is not of
        this form cannot be standardised on the device."""
        raise NotImplementedError(f"{self.__class__.__name__} has no device standardisation table")

    def standardisation_table(self, input_feature_names: List[str]) -> Tuple[List[int], List[float], List[float]]:
        amap = self.affine_map()
        kinds, subs, divs = [], [], []
        for name in input_feature_names:
            kind, sub, div = amap[name]          # KeyError for an unknown column, like `forward`
            kinds.append(int(kind)); subs.append(float(sub)); divs.append(float(div))
        return kinds, subs, divs


class IceCube86(Detector):
    """xyz/500, (t-1e4)/3e4, log10(charge), (rde-1.25)/0.25, pmt_area/0.05, hlc unchanged (icecube.py:21-48)."""

    def affine_map(self) -> Dict[str, Tuple[int, float, float]]:
        return {
            "dom_x": (STD_AFFINE, 0.0, 500.0), "dom_y": (STD_AFFINE, 0.0, 500.0), "dom_z": (STD_AFFINE, 0.0, 500.0),
            "dom_time": (STD_AFFINE, 1.0e04, 3.0e4), "charge": (STD_LOG10, 0.0, 1.0),
            "rde": (STD_AFFINE, 1.25, 0.25), "pmt_area": (STD_AFFINE, 0.0, 0.05), "hlc": (STD_IDENTITY, 0.0, 1.0),
        }

    def feature_map(self) -> Dict[str, Callable]:
        return {
            "dom_x": lambda v: v / 500.0, "dom_y": lambda v: v / 500.0, "dom_z": lambda v: v / 500.0,
            "dom_time": lambda v: (v - 1.0e04) / 3.0e4, "charge": lambda v: torch.log10(v),
            "rde": lambda v: (v - 1.25) / 0.25, "pmt_area": lambda v: v / 0.05, "hlc": lambda v: v,
        }


class ORCA150SuperDense(Detector):
    """Prometheus ORCA150SuperDense: xy/100, (z+350)/100, t/1.05e4 (reference detector/prometheus.py:11-39)."""

    def affine_map(self) -> Dict[str, Tuple[int, float, float]]:
        return {"sensor_pos_x": (STD_AFFINE, 0.0, 100.0), "sensor_pos_y": (STD_AFFINE, 0.0, 100.0),
                "sensor_pos_z": (STD_AFFINE, -350.0, 100.0), "t": (STD_AFFINE, 0.0, 1.05e04)}

    def feature_map(self) -> Dict[str, Callable]:
        return {"sensor_pos_x": lambda v: v / 100, "sensor_pos_y": lambda v: v / 100,
                "sensor_pos_z": lambda v: (v + 350) / 100, "t": lambda v: v / 1.05e04}


class Prometheus(ORCA150SuperDense):
    """Reference to ORCA150SuperDense (detector/prometheus.py:365-366): the detector of BASELINE configs[0]."""


class IdentityDetector(Detector):
    """No standardisation at all (inputs are already standardised; not a reference class)."""

    def feature_map(self) -> Dict[str, Callable]:
        return {}

    def forward(self, input_features: torch.Tensor, input_feature_names: List[str]) -> torch.Tensor:
        return input_features.clone()

    def standardisation_table(self, input_feature_names: List[str]) -> Tuple[List[int], List[float], List[float]]:
        n = len(input_feature_names)
        return [STD_IDENTITY] * n, [0.0] * n, [1.0] * n
