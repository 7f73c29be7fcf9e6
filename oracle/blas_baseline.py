"""BLAS-speed CPU statement of the 1:N search, for TIMING the reference-side cost fairly.  TEST INFRASTRUCTURE ONLY.

The reference computes embedding distances with numpy / TensorFlow-CPU library calls (api.py:103, predictions.py:126,
common/losses.py:40 `tf.matmul`), i.e. a multithreaded sgemm, not a hand-ordered reduction.  This module is that
path for the gallery workload: normalise, sgemm in row chunks, per-query top-k (torch.topk, the fastest CPU top-k
probed in SURVEY.md section 6).  Its scores differ from the canonical oracle in the last bits, so it is used for the
`cpu_baseline` / `--impl reference` timings only - correctness is always judged against oracle/dif_oracle.c.
"""
from __future__ import annotations

import numpy as np


def gallery_search_blas(gallery_unit: np.ndarray, queries_unit: np.ndarray, k: int, chunk: int = 262144):
    """gallery_unit [N, D], queries_unit [Q, D] already L2-normalised fp32.  Returns (scores [Q,k], rows [Q,k])."""
    import torch

    g = torch.from_numpy(np.ascontiguousarray(gallery_unit, dtype=np.float32))
    q = torch.from_numpy(np.ascontiguousarray(queries_unit, dtype=np.float32))
    best_s = torch.full((q.shape[0], k), -float("inf"))
    best_r = torch.full((q.shape[0], k), -1, dtype=torch.int64)
    for lo in range(0, g.shape[0], chunk):
        s = q @ g[lo:lo + chunk].T
        ts, tr = torch.topk(s, min(k, s.shape[1]), dim=1)
        cat_s = torch.cat([best_s, ts], dim=1)
        cat_r = torch.cat([best_r, tr + lo], dim=1)
        best_s, idx = torch.topk(cat_s, k, dim=1)
        best_r = torch.gather(cat_r, 1, idx)
    return best_s.numpy(), best_r.numpy()


def num_threads() -> int:
    import torch

    return torch.get_num_threads()
