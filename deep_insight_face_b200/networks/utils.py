"""Scalar helpers mirroring the names in deep_insight_face/networks/utils.py:4-39 (host-side, as in the reference).

    distance                       squared L2 of two embeddings                       (utils.py:4-9)
    distance_to_proba              d -> 1 / (1 + d)                                   (utils.py:12-17)
    gaussian_kernel_dist_to_prob   d -> exp(-d / (2 * tuning_factor ** 2))            (utils.py:20-29)
    calc_mean_score                mean of the ranks 1..10 under a normalised score   (utils.py:32-39)
"""
from __future__ import annotations

import numpy as np

__all__ = ["distance", "distance_to_proba", "gaussian_kernel_dist_to_prob", "calc_mean_score"]


def distance(emb1, emb2):
    delta = np.asarray(emb1) - np.asarray(emb2)
    return np.sum(delta * delta)


def distance_to_proba(distance):
    return 1.0 / (1.0 + distance)


def gaussian_kernel_dist_to_prob(distance, tuning_factor=1.0):
    two_sigma_sq = 2.0 * tuning_factor * tuning_factor
    return np.exp(-distance / two_sigma_sq)


def calc_mean_score(score_dist):
    weights = np.asarray(score_dist, dtype=np.float64)
    weights = weights / weights.sum()
    ranks = np.arange(1, weights.shape[0] + 1)
    return float((weights * ranks).sum()) if weights.shape[0] == 10 else (weights * np.arange(1, 11)).sum()
