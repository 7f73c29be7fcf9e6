"""ctypes binding of oracle/dif_oracle.c (canonical fp32 reductions, 1:N search, pair distance, sweep).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Pair distance / threshold counts are pinned against
the live reference functions by tests/golden (see tests/golden/make_golden.py).  The reference has no 1:N
routine; the search is pinned as "rank the gallery by the reference's own distance function"
(evaluation/utility.py:52-66 run over every query x row pair by tests/golden/make_golden_gallery.py:
same rows, same order, same numbers); its tie rule (lower row first) and the id / empty-slot conventions are
this build's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dif_oracle.c")
LIB = os.path.join(HERE, "libdif_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-mfma", "-mavx2", "-fPIC", "-shared",
               "-o", LIB, SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
        L.dif_or_canon_dot.restype = f32
        L.dif_or_canon_dot.argtypes = [vp, vp, i32]
        L.dif_or_canon_sqdist.restype = f32
        L.dif_or_canon_sqdist.argtypes = [vp, vp, i32]
        L.dif_or_normalize_rows.argtypes = [vp, i64, i32, vp]
        L.dif_or_row_sqnorm.argtypes = [vp, i64, i32, vp]
        L.dif_or_cross.argtypes = [vp, i64, vp, i64, i32, i32, vp]
        L.dif_or_synth_value.restype = f32
        L.dif_or_synth_value.argtypes = [u64, u64, u64, u64]
        L.dif_or_synth_rows.argtypes = [u64, i64, i64, i32, vp]
        L.dif_or_gallery_search.argtypes = [vp, i64, i32, i32, vp, i32, i32, vp, vp]
        L.dif_or_topk_merge.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        L.dif_or_pair_distance.argtypes = [vp, vp, i64, i32, i32, vp]
        L.dif_or_threshold_sweep.argtypes = [vp, vp, vp, i64, vp, i32, vp]
        L.dif_or_num_threads.restype = i32
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().dif_or_num_threads())


def synth_rows(seed: int, row0: int, n: int, D: int) -> np.ndarray:
    out = np.empty((n, D), dtype=np.float32)
    lib().dif_or_synth_rows(seed, row0, n, D, out.ctypes.data)
    return out


def normalize_rows(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty_like(x)
    lib().dif_or_normalize_rows(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def row_sqnorm(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().dif_or_row_sqnorm(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def cross(a, b, metric: int) -> np.ndarray:
    """out[i, j] = canonical dot (metric 1) or squared distance (metric 0) of a_i and b_j."""
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().dif_or_cross(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], a.shape[1], metric, out.ctypes.data)
    return out


def gallery_search(gallery, queries, k: int, metric: int, normalize: bool | None = None):
    """Oracle of Gallery.search: rows are normalised first for cosine (as the library stores them)."""
    g, q = _f32(gallery), _f32(queries)
    if normalize is None:
        normalize = metric == 1
    if normalize:
        g, q = normalize_rows(g), normalize_rows(q)
    scores = np.empty((q.shape[0], k), dtype=np.float32)
    rows = np.empty((q.shape[0], k), dtype=np.int64)
    lib().dif_or_gallery_search(g.ctypes.data, g.shape[0], g.shape[1], metric, q.ctypes.data, q.shape[0], k,
                                scores.ctypes.data, rows.ctypes.data)
    return scores, rows


def topk_merge(scores, grows, metric: int):
    """scores/grows [world, Q, k] -> merged [Q, k]."""
    s, r = _f32(scores), np.ascontiguousarray(grows, dtype=np.int64)
    world, Q, k = s.shape
    out_s = np.empty((Q, k), dtype=np.float32)
    out_r = np.empty((Q, k), dtype=np.int64)
    lib().dif_or_topk_merge(s.ctypes.data, r.ctypes.data, world, Q, k, metric, out_s.ctypes.data, out_r.ctypes.data)
    return out_s, out_r


def pair_distance(e1, e2, metric: int) -> np.ndarray:
    e1, e2 = _f32(e1), _f32(e2)
    out = np.empty(e1.shape[0], dtype=np.float32)
    lib().dif_or_pair_distance(e1.ctypes.data, e2.ctypes.data, e1.shape[0], e1.shape[1], metric, out.ctypes.data)
    return out


def threshold_sweep(dist, issame, thresholds, select=None) -> np.ndarray:
    d = _f32(dist)
    s = np.ascontiguousarray(issame, dtype=np.uint8)
    t = np.ascontiguousarray(np.atleast_1d(np.asarray(thresholds, dtype=np.float64)))
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.uint8)
    out = np.empty((t.shape[0], 4), dtype=np.int64)
    lib().dif_or_threshold_sweep(d.ctypes.data, s.ctypes.data, None if sel is None else sel.ctypes.data, d.shape[0],
                                 t.ctypes.data, t.shape[0], out.ctypes.data)
    return out
