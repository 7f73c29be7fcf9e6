"""Property tests (hypothesis) of the host-side logic and the oracle: fold layout against sklearn's KFold, shard
partitioning, merge invariants, threshold counts against the reference's own boolean reductions."""
import numpy as np
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(n=st.integers(10, 700), k=st.integers(2, 10))
def test_kfold_ids_match_sklearn(n, k):
    from sklearn.model_selection import KFold

    from deep_insight_face_b200.evaluation.utility import kfold_ids

    ids = kfold_ids(n, k)
    for f, (_, test) in enumerate(KFold(n_splits=k, shuffle=False).split(np.arange(n))):
        assert np.array_equal(np.flatnonzero(ids == f), test)


@settings(max_examples=100, deadline=None)
@given(n=st.integers(0, 10**9), w=st.integers(1, 64))
def test_shard_ranges(n, w):
    from deep_insight_face_b200.gallery import shard_range

    spans = [shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10**6), world=st.integers(1, 5), k=st.integers(1, 12), metric=st.sampled_from([0, 1]))
def test_merge_of_shard_results_is_the_global_topk(orc, seed, world, k, metric):
    rng = np.random.default_rng(seed)
    N, Q, D = int(rng.integers(1, 400)), 6, 32
    rows = np.round(rng.standard_normal((N, D)), 1).astype(np.float32)   # coarse values -> plenty of exact ties
    rows[rng.integers(0, N, size=N // 3)] = rows[0]
    q = np.round(rng.standard_normal((Q, D)), 1).astype(np.float32)
    s, r = orc.gallery_search(rows, q, k, metric, normalize=False)
    cuts = np.sort(rng.integers(0, N + 1, size=world - 1)) if world > 1 else np.array([], dtype=int)
    bounds = [0, *cuts.tolist(), N]
    parts = []
    for lo, hi in zip(bounds, bounds[1:]):
        if hi > lo:
            ss, rr = orc.gallery_search(rows[lo:hi], q, k, metric, normalize=False)
        else:
            ss, rr = np.zeros((Q, k), np.float32), np.full((Q, k), -1, np.int64)
        parts.append((ss, np.where(rr >= 0, rr + lo, -1)))
    ms, mr = orc.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), metric)
    assert np.array_equal(mr, r) and np.array_equal(ms.view(np.uint32), s.view(np.uint32))
    # ties resolve to the lower row: within equal scores rows ascend
    for row_s, row_r in zip(s, r):
        v = row_r >= 0
        same = (np.diff(row_s[v]) == 0)
        assert np.all(np.diff(row_r[v])[same] > 0)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10**6), n=st.integers(1, 300), t=st.integers(1, 50))
def test_threshold_counts_equal_the_reference_formulas(orc, seed, n, t):
    rng = np.random.default_rng(seed)
    dist = np.round(rng.random(n) * 4, 2).astype(np.float32)          # values ON thresholds: strict '<' matters
    same = rng.random(n) > 0.5
    thr = np.sort(np.round(rng.random(t) * 4, 2))
    counts = orc.threshold_sweep(dist, same, thr)
    for i, th in enumerate(thr):
        pred = np.less(dist, th)                                        # evaluation/utility.py:37
        want = (np.sum(pred & same), np.sum(pred & ~same), np.sum(~pred & ~same), np.sum(~pred & same))
        assert tuple(int(c) for c in counts[i]) == tuple(int(w) for w in want)
