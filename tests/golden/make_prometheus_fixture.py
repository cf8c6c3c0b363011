"""Fixture for BASELINE configs[0]: the reference's example data base `data/examples/sqlite/prometheus/prometheus-events.db`
(50 events, 1 872 pulses, pulsemap `total`, truth table `mc_truth`; examples/04_training/01_train_dynedge.py:195-213) as
plain arrays, plus the standardised features produced by the REFERENCE's own, unmodified `detector/prometheus.py`
(`Prometheus` = `ORCA150SuperDense`, :11-39, :365) through its own `Detector._standardize` (detector.py:63-77).

The GPU box has no /root/reference, so the 1 872 x 4 pulse table travels as tests/golden/prometheus_events.npz (37 KB).
Run (only in the build container): python tests/golden/make_prometheus_fixture.py
"""
from __future__ import annotations

import importlib
import os
import sqlite3
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

DB = "/root/reference/data/examples/sqlite/prometheus/prometheus-events.db"
FEATURES = ["sensor_pos_x", "sensor_pos_y", "sensor_pos_z", "t"]        # FEATURES.PROMETHEUS, data/constants.py:26-31


def load_reference_prometheus():
    mg.install_shims()
    mg._mod("graphnet.utilities")
    mg._mod("graphnet.utilities.decorators", final=lambda f: f)
    mg._mod("graphnet.constants", PROMETHEUS_GEOMETRY_TABLE_DIR="/nonexistent", ICECUBE_GEOMETRY_TABLE_DIR="/nonexistent")
    m = mg._mod("graphnet.models.detector")
    m.__path__ = [os.path.join(mg.REF_SRC, "graphnet", "models", "detector")]
    return importlib.import_module("graphnet.models.detector.prometheus").Prometheus


def main() -> None:
    con = sqlite3.connect(f"file:{DB}?mode=ro", uri=True)
    # SQLiteDataset queries one event at a time in table order (dataset/sqlite/sqlite_dataset.py); events in event_no order
    rows = con.execute(f"select event_no, {', '.join(FEATURES)} from total order by event_no, rowid").fetchall()
    truth = dict(con.execute("select event_no, total_energy from mc_truth").fetchall())
    arr = np.asarray(rows, dtype=np.float64)
    event_no = arr[:, 0].astype(np.int64)
    raw = arr[:, 1:].astype(np.float32)
    uniq, counts = np.unique(event_no, return_counts=True)
    energy = np.asarray([truth[int(e)] for e in uniq], dtype=np.float32)
    cls = load_reference_prometheus()
    det = cls.__new__(cls)
    torch.nn.Module.__init__(det)
    std = det._standardize(torch.from_numpy(raw.copy()), FEATURES).numpy()
    np.savez_compressed(os.path.join(HERE, "prometheus_events.npz"), raw=raw, standardized=std, event_no=uniq,
                        n_pulses=counts.astype(np.int32), total_energy=energy, features=np.asarray(FEATURES))
    dup = sum(len(raw[event_no == e]) - len(np.unique(raw[event_no == e][:, :3], axis=0)) for e in uniq)
    print("events", len(uniq), "pulses", len(raw), "min/max pulses", counts.min(), counts.max(),
          "events < 9 pulses", int((counts < 9).sum()), "duplicate-xyz pulses", dup)


if __name__ == "__main__":
    main()
