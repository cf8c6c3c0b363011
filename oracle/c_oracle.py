"""ctypes loader for the plain-C oracle (oracle/oracle_c.c). TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib: Optional[ctypes.CDLL] = None


def build() -> str:
    src = os.path.join(_HERE, "oracle_c.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_knn_table.restype = ctypes.c_int64
        _lib.oracle_knn_edge_index.restype = ctypes.c_int64
    return _lib


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def knn_table(x: np.ndarray, cols: Sequence[int], ptr: np.ndarray, k: int,
              threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Neighbour table [N, k+1] (int32, -1 padded) and degrees [N]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    cols_a = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
    n = x.shape[0]
    nbr = np.empty((n, k + 1), dtype=np.int32)
    deg = np.empty((n,), dtype=np.int32)
    nseg = len(ptr) - 1
    fn = lib().oracle_knn_table

    def run(lo: int, hi: int) -> int:
        return fn(_p(x), ctypes.c_int64(x.shape[1]), _p(cols_a), ctypes.c_int32(len(cols_a)), _p(ptr),
                  ctypes.c_int64(lo), ctypes.c_int64(hi), ctypes.c_int32(k), _p(nbr), _p(deg))

    if threads <= 1 or nseg < 2 * threads:
        assert run(0, nseg) >= 0
    else:  # ctypes releases the GIL: balance segment ranges by sum n^2
        from concurrent.futures import ThreadPoolExecutor
        cost = np.cumsum((ptr[1:] - ptr[:-1]).astype(np.float64) ** 2)
        cuts = [0] + [int(np.searchsorted(cost, cost[-1] * (i + 1) / (4 * threads))) + 1
                      for i in range(4 * threads - 1)] + [nseg]
        cuts = sorted(set(min(c, nseg) for c in cuts))
        with ThreadPoolExecutor(threads) as ex:
            for r in ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])):
                assert r >= 0
    return nbr, deg


def knn_edge_index(x: np.ndarray, cols: Sequence[int], ptr: np.ndarray, k: int, threads: int = 1) -> np.ndarray:
    nbr, deg = knn_table(x, cols, ptr, k, threads)
    e = int(deg.sum())
    out = np.empty((2, e), dtype=np.int64)
    got = lib().oracle_knn_edge_index(_p(nbr), _p(deg), ctypes.c_int64(nbr.shape[0]), ctypes.c_int32(k),
                                      _p(out[0]), _p(out[1]))
    assert got == e
    return out


def homophily(x: np.ndarray, col: int, edge_index: np.ndarray, batch: np.ndarray, nseg: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    src = np.ascontiguousarray(edge_index[0], dtype=np.int64)
    dst = np.ascontiguousarray(edge_index[1], dtype=np.int64)
    batch = np.ascontiguousarray(batch, dtype=np.int64)
    out = np.empty((nseg,), dtype=np.float32)
    lib().oracle_homophily(_p(x), ctypes.c_int64(x.shape[1]), ctypes.c_int32(col), _p(src), _p(dst),
                           ctypes.c_int64(src.shape[0]), _p(batch), ctypes.c_int64(nseg), _p(out))
    return out


_SCHEMES = {"min": 0, "max": 1, "sum": 2, "mean": 3}


def segment_pool(x: np.ndarray, ptr: np.ndarray, scheme: str) -> Tuple[np.ndarray, np.ndarray]:
    x = np.ascontiguousarray(x, dtype=np.float32)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    nseg, c = len(ptr) - 1, x.shape[1]
    out = np.empty((nseg, c), dtype=np.float32)
    arg = np.full((nseg, c), -1, dtype=np.int64)
    lib().oracle_segment_pool(_p(x), ctypes.c_int64(c), ctypes.c_int64(c), _p(ptr), ctypes.c_int64(nseg),
                              ctypes.c_int32(_SCHEMES[scheme]), _p(out), _p(arg))
    return out, arg
