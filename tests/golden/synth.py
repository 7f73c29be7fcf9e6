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




def gallery_case(seed, N, Q, D, noise=0.3):
    """A 1:N probe set: N(0, 1) gallery rows (not normalised) and Q queries, each a gallery row plus noise."""
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((N, D)).astype(np.float32)
    pick = rng.integers(0, N, size=Q)
    q = (rows[pick] + np.float32(noise) * rng.standard_normal((Q, D)).astype(np.float32)).astype(np.float32)
    return rows, q, pick


GALLERY_CASES = [  # name, seed, N, Q, D, k        (make_golden_gallery.py / the 1:N pins of the tests)
    ("small", 61, 3000, 40, 128, 10),
    ("wide", 62, 1500, 24, 512, 5),
    ("short", 63, 300, 9, 96, 24),
]


def check_ranking(ref_dist, ref_rows, scores, rows, metric, gap=2e-6, tol=2e-6):
    """A search result (scores / rows [Q, k], best first) against the reference's k + 1 smallest distances
    (make_golden_gallery.py).  Every rank whose reference distance is separated from both neighbours by more than
    `gap` (relative) must hold the same gallery row; every score must be the reference distance (metric 0: the
    squared distance itself; metric 1: cos(pi * d) is the similarity the search reports).  Returns the share of
    ranks that were decidable, so callers can assert the check was not vacuous."""
    Q, k = rows.shape
    d = ref_dist
    scale = np.maximum(np.abs(d[:, :k]), 1.0)
    up = (d[:, 1 : k + 1] - d[:, :k]) > gap * scale
    down = np.concatenate([np.ones((Q, 1), dtype=bool), up[:, :-1]], axis=1)
    decidable = up & down
    assert np.array_equal(rows[decidable], ref_rows[:, :k][decidable]), "ranking differs from the reference's distance order"
    want = d[:, :k] if metric == 0 else np.cos(np.pi * d[:, :k])
    err = np.abs(scores.astype(np.float64) - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() <= tol, f"scores differ from the reference's distances: {err.max():.3g}"
    return float(decidable.mean())
