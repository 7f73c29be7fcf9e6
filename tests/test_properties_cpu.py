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


def test_bit_reversed_counter_tree_is_the_canonical_butterfly():
    """csrc/canon_mm.cuh lets ONE thread fold the 32 chain sums of the canonical dot product: it visits the chains in
    bit-reversed order and keeps a binary counter of pending partial sums.  That must be, addition for addition, the
    butterfly t[i] += t[i ^ o], o = 16 .. 1, of dif_canon.cuh / oracle/dif_oracle.c (fp32 addition is commutative, not
    associative: the PAIRING is what has to match).  Checked here in numpy float32 on adversarial magnitudes; the
    half-chain variant of the cluster kernel (even chains -> y0, odd chains -> y1, total y0 + y1) as well."""
    rng = np.random.default_rng(0)
    f32 = np.float32
    for trial in range(200):
        t = (rng.standard_normal(32) * 10.0 ** rng.integers(-6, 7, size=32)).astype(f32)
        bt = t.copy()
        for o in (16, 8, 4, 2, 1):
            bt = (bt + bt[np.arange(32) ^ o]).astype(f32)
        want = bt[0]
        assert all(bt[i] == want or (np.isnan(bt[i]) and np.isnan(want)) for i in range(32))   # every lane agrees

        def counter_tree(order):
            stack = {}
            for n, chain in enumerate(order):
                v = t[chain]
                lvl, m = 0, n
                while m & 1:
                    v = f32(stack[lvl] + v)
                    m >>= 1
                    lvl += 1
                stack[lvl] = v
            return stack[lvl]

        bitrev = [int(format(n, "05b")[::-1], 2) for n in range(32)]
        assert counter_tree(bitrev).tobytes() == want.tobytes()
        y0, y1 = counter_tree(bitrev[:16]), counter_tree(bitrev[16:])
        assert all(c % 2 == 0 for c in bitrev[:16]) and all(c % 2 == 1 for c in bitrev[16:])
        assert f32(y0 + y1).tobytes() == want.tobytes()
