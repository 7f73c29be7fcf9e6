"""deep_insight_face/networks/utils.py:4-39: scalar distance helpers (trivial, host-side as in the reference)."""
import numpy as np


def distance(emb1, emb2):
    """utils.py:4-9: squared L2 of two embeddings."""
    return np.sum(np.square(np.asarray(emb1) - np.asarray(emb2)))


def distance_to_proba(distance):
    """utils.py:12-17."""
    return 1 / (1 + distance)


def gaussian_kernel_dist_to_prob(distance, tuning_factor=1.0):
    """utils.py:20-29."""
    return np.exp(-distance / (2 * tuning_factor**2))


def calc_mean_score(score_dist):
    """utils.py:32-39."""
    score_dist = np.array(score_dist)
    score_dist = score_dist / score_dist.sum()
    return (score_dist * np.arange(1, 11)).sum()
