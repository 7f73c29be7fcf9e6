"""Entry points no other GPU test reaches: the single-pair twin get_emd_distance (evaluation/utility.py:174-188)
and Gallery.reset (dif_gallery_reset).  Written after the round's GPU budget was spent; first run is the driver's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_get_emd_distance_is_the_single_pair_twin(gpu, golden):
    from synth import pairs

    from deep_insight_face_b200.evaluation import utility as U

    emb, _ = pairs(4, 600, 128)
    e1, e2 = emb[0::2], emb[1::2]
    for i in (0, 7, 599):      # the reference's call form (evals.py:119): two 1-D embeddings, metric 0 -> a scalar
        want = np.sum(np.square(np.subtract(e1[i].astype(np.float64), e2[i].astype(np.float64))), 0)
        got = U.get_emd_distance(e1[i], e2[i], 0)
        assert np.ndim(got) == 0 and abs(float(got) - want) <= 1e-6 * max(1.0, want)
    np.testing.assert_allclose(U.get_emd_distance(e1, e2, 1), golden["small_dist1"], rtol=1e-4, atol=1e-6)
    with pytest.raises(RuntimeError):
        U.get_emd_distance(e1, e2, 2)


def test_reset_empties_the_gallery_and_it_fills_again(gpu, orc):
    from deep_insight_face_b200.gallery import Gallery

    first, second = orc.synth_rows(5, 0, 3000, 128), orc.synth_rows(6, 0, 1000, 128)
    pick = np.random.default_rng(1).integers(0, 1000, size=40)
    q = second[pick] + 0.3 * orc.synth_rows(36, 0, 40, 128)
    with Gallery(4000, 128, "cosine", "tf32x3") as g:
        g.add(first, np.arange(3000, dtype=np.int64) * 3 + 7)
        g.search(q, 10)
        g.reset()
        assert len(g) == 0
        g.add(second)                      # ids are the row numbers again, no explicit ids required after a reset
        assert len(g) == 1000
        s, ids, r = g.search(q, 10, return_rows=True)
        ws, wr = orc.gallery_search(second, q, 10, 1)
        assert np.array_equal(r.astype(np.int64), wr) and np.array_equal(s.view(np.uint32), ws.view(np.uint32))
        assert np.array_equal(np.asarray(ids).astype(np.int64), wr)


@pytest.mark.parametrize("name,seed,noise", [("hard10", 5, 1.0), ("hard15", 6, 1.5)])
def test_val_at_far_matches_the_reference(gpu, golden, name, seed, noise):
    """calculate_val / evaluate (utility.py:80-119, :10-33) on overlapping classes against the reference's own run
    (make_golden.py; its interp1d call repaired for repeated x, named in the keys).  Distances here are the canonical
    fp32 ones, the reference's are numpy's: a pair within an ulp of the selected threshold may change sides, which
    moves VAL or FAR of one fold by 1/60 - hence 2 pairs of 600 on the means."""
    from synth import pairs

    from deep_insight_face_b200.evaluation import utility as U

    emb, issame = pairs(seed, 1200, 64, noise=noise)
    e1, e2 = emb[0::2], emb[1::2]
    thr = np.arange(0, 4, 0.001)
    for metric in (0, 1):
        for sm in (False, True):
            for far_target in (1e-3, 1e-2):
                val, std, far = U.calculate_val(thr, e1, e2, issame, far_target, 10, metric, subtract_mean=sm)
                want = golden[f"{name}_val_repaired{metric}_{int(sm)}_far{far_target:g}"]
                assert abs(val - want[0]) <= 2.0 / 600 and abs(far - want[2]) <= 2.0 / 600, (metric, sm, far_target, val, far, want)
                assert abs(std - want[1]) <= 1e-2
    out = U.evaluate(emb, issame)
    want = golden[f"{name}_evaluate_repaired"]
    assert abs(np.mean(out[2]) - want[0]) <= 2.0 / 120 and abs(out[4] - want[2]) <= 2.0 / 600 and abs(out[6] - want[4]) <= 2.0 / 600


def test_verify_methods_match_the_reference(gpu, capsys):
    """TripletPrediction.verify / SiamesePrediction.verify against the reference's own methods executed on a stub self
    (tests/golden/make_golden_host.py): distance, decision and printed line."""
    import json
    import os

    from make_golden_host import verify_cases

    from deep_insight_face_b200.predictions import SiamesePrediction, TripletPrediction

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_reference.json")) as f:
        ref = {r["name"]: r for r in json.load(f)["verify"]}
    for name, enc, stored, thr in verify_cases():
        for key, obj, db in (("triplet", TripletPrediction(), {"who": stored[0]}), ("siamese", SiamesePrediction(), {"who": stored})):
            capsys.readouterr()
            dist, ok = obj.verify(enc, "who", db, threshold=thr)
            want = ref[name][key]
            assert capsys.readouterr().out == want["printed"] and ok is want["is_valid"], (name, key)
            assert abs(dist - want["dist"]) <= 1e-5 * max(want["dist"], 1e-3), (name, key, dist, want["dist"])
