"""Seeded synthetic verification pairs shared by make_golden.py and the tests (numpy only)."""
import numpy as np


def pairs(seed, n_pairs, D, noise=0.5):
    """Interleaved embeddings as utility.py:18-19 expects: even rows = first of pair; same = shared centroid."""
    rng = np.random.default_rng(seed)
    issame = np.tile(np.repeat([True, False], 30), n_pairs // 60 + 1)[:n_pairs]
    c = rng.standard_normal((n_pairs, D)).astype(np.float32)
    e1 = c + noise * rng.standard_normal((n_pairs, D)).astype(np.float32)
    other = rng.standard_normal((n_pairs, D)).astype(np.float32)
    e2 = np.where(issame[:, None], c, other) + noise * rng.standard_normal((n_pairs, D)).astype(np.float32)
    n = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    e1, e2 = n(e1), n(e2)
    emb = np.empty((2 * n_pairs, D), dtype=np.float32)
    emb[0::2], emb[1::2] = e1, e2
    return emb, issame


