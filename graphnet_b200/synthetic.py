"""Synthetic IceCube-like pulse maps for benchmarks and full-size tests.

Per event: n = clip(floor(LogNormal(ln 100, sigma)), 2, n_max) pulses on ceil(0.6 n) distinct DOMs
chosen around a random vertex (p ~ exp(-r / 150 m)) of an 86-string x 60-DOM hexagonal detector
(125 m string spacing, 17 m DOM spacing -- the survey's fallback geometry, used unconditionally so the
GPU box needs no reference file); pulses are assigned to those DOMs with replacement, which yields the
20-40 % duplicate-xyz pulses of real data. Features follow FEATURES.ICECUBE86 order
(dom_x, dom_y, dom_z, dom_time, charge, rde, pmt_area) and are standardised with the IceCube86
constants (reference: src/graphnet/models/detector/icecube.py:21-48).
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

FEATURES_ICECUBE86 = ["dom_x", "dom_y", "dom_z", "dom_time", "charge", "rde", "pmt_area"]
_GEOMETRY: Optional[np.ndarray] = None


def detector_geometry() -> np.ndarray:
    """[5160, 4]: x, y, z (m), rde."""
    global _GEOMETRY
    if _GEOMETRY is None:
        pts = []
        # hexagonal spiral of 86 strings
        coords = [(0, 0)]
        ring = 1
        while len(coords) < 86:
            q, r = ring, 0
            for dq, dr in [(-1, 1), (-1, 0), (0, -1), (1, -1), (1, 0), (0, 1)]:
                for _ in range(ring):
                    if len(coords) < 86:
                        coords.append((q, r))
                    q, r = q + dq, r + dr
            ring += 1
        for si, (q, r) in enumerate(coords):
            sx = 125.0 * (q + 0.5 * r)
            sy = 125.0 * (np.sqrt(3.0) / 2.0) * r
            for d in range(60):
                rde = 1.35 if (si >= 78 and d >= 10) else 1.0
                pts.append((sx, sy, 500.0 - 17.0 * d, rde))
        _GEOMETRY = np.asarray(pts, dtype=np.float64)
    return _GEOMETRY


def event_sizes(num_events: int, rng: np.random.Generator, sigma: float = 1.0, n_max: int = 5000,
                median: float = 100.0) -> np.ndarray:
    n = np.floor(rng.lognormal(mean=np.log(median), sigma=sigma, size=num_events))
    return np.clip(n, 2, n_max).astype(np.int64)


def make_batch(num_events: int, seed: int = 20240607, sigma: float = 1.0, n_max: int = 5000,
               sizes: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Returns x [N,7] fp32 (standardised), batch [N] i64, ptr [B+1] i64, n_pulses [B] i32 and labels."""
    rng = np.random.default_rng(seed)
    geo = detector_geometry()
    if sizes is None:
        sizes = event_sizes(num_events, rng, sigma, n_max)
    sizes = np.asarray(sizes, dtype=np.int64)
    total = int(sizes.sum())
    x = np.empty((total, 7), dtype=np.float64)
    pos = 0
    for n in sizes:
        n = int(n)
        vertex = np.array([rng.uniform(-400, 400), rng.uniform(-400, 400), rng.uniform(-400, 400)])
        r = np.linalg.norm(geo[:, :3] - vertex, axis=1)
        w = np.exp(-r / 150.0)
        w /= w.sum()
        ndom = min(int(np.ceil(0.6 * n)), geo.shape[0])
        doms = rng.choice(geo.shape[0], size=ndom, replace=False, p=w)
        pick = doms[rng.integers(0, ndom, size=n)]
        x[pos:pos + n, 0:3] = geo[pick, :3]
        x[pos:pos + n, 3] = rng.normal(1.0e4, 1.5e3, size=n)
        x[pos:pos + n, 4] = rng.lognormal(0.0, 0.7, size=n)
        x[pos:pos + n, 5] = geo[pick, 3]
        x[pos:pos + n, 6] = 0.0444
        pos += n
    # IceCube86 standardisation
    x[:, 0:3] /= 500.0
    x[:, 3] = (x[:, 3] - 1.0e4) / 3.0e4
    x[:, 4] = np.log10(x[:, 4])
    x[:, 5] = (x[:, 5] - 1.25) / 0.25
    x[:, 6] /= 0.05
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    batch = np.repeat(np.arange(len(sizes), dtype=np.int64), sizes)
    direction = rng.normal(size=(len(sizes), 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    return {
        "x": x.astype(np.float32), "batch": batch, "ptr": ptr, "n_pulses": sizes.astype(np.int32),
        "energy": (10.0 ** rng.uniform(0.0, 4.0, size=len(sizes))).astype(np.float32),
        "direction": direction.astype(np.float32),
    }
