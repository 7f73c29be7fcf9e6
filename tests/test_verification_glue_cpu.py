"""The host logic of calculate_val / evaluate (evaluation/utility.py mirror of the reference's utility.py:10-33,80-119)
on the CPU against a FAKE library: dif_pair_distance_host hands back the REFERENCE's own distances (tests/golden) and
dif_threshold_sweep_host counts with numpy through the raw pointers it is given.  Fed the reference's distances, the
k-fold VAL @ FAR selection must then reproduce the numbers the reference's calculate_val / evaluate produced
(make_golden.py; one dependency call of the reference repaired for that run - interp1d with repeated x - and named in
the keys), float for float: fold layout, train = all - test counts, the far_train -> threshold interpolation, the
0-threshold branch, mean / std over folds."""
import ctypes as C

import numpy as np
import pytest

from synth import pairs


def _view(ptr, shape, ctype):
    return np.ctypeslib.as_array(C.cast(int(ptr), C.POINTER(ctype)), shape=tuple(shape))


class FakeLib:
    def __init__(self, distances):
        self.distances = distances      # metric -> the reference's distance vector for the pair set under test
        self.sweeps = 0

    def dif_pair_distance_host(self, e1, e2, n, D, metric, mean, out):
        assert not mean
        _view(out, (n,), C.c_float)[:] = self.distances[metric][:n]
        return 0

    def dif_threshold_sweep_host(self, dist, issame, fold, n, n_folds, thr, T, counts):
        d = _view(dist, (n,), C.c_float).astype(np.float64)
        same = _view(issame, (n,), C.c_uint8).astype(bool)
        f = _view(fold, (n,), C.c_int32) if fold else np.zeros(n, np.int32)
        t = _view(thr, (T,), C.c_double)
        out = _view(counts, (n_folds, T, 4), C.c_int64)
        for k in range(n_folds):
            sel = f == k
            pred = d[sel][None, :] < t[:, None]                      # np.less against the float64 thresholds
            s = same[sel][None, :]
            out[k, :, 0] = (pred & s).sum(1)
            out[k, :, 1] = (pred & ~s).sum(1)
            out[k, :, 2] = (~pred & ~s).sum(1)
            out[k, :, 3] = (~pred & s).sum(1)
        self.sweeps += 1
        return 0


@pytest.mark.parametrize("name,seed,noise", [("hard10", 5, 1.0), ("hard15", 6, 1.5)])
def test_val_at_far_reproduces_the_reference_on_reference_distances(monkeypatch, golden, name, seed, noise):
    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.evaluation import utility as U

    emb, issame = pairs(seed, 1200, 64, noise=noise)
    e1, e2 = emb[0::2], emb[1::2]
    lib = FakeLib({m: golden[f"{name}_dist{m}"].astype(np.float32) for m in (0, 1)})
    monkeypatch.setattr(_ffi, "load_library", lambda: lib)
    monkeypatch.setattr(_ffi, "init", lambda device=None: None)
    thr = np.arange(0, 4, 0.001)
    for metric in (0, 1):
        for far_target in (1e-3, 1e-2):
            got = U.calculate_val(thr, e1, e2, issame, far_target, nrof_folds=10, distance_metric=metric)
            want = golden[f"{name}_val_repaired{metric}_0_far{far_target:g}"]
            assert np.array_equal(np.array(got), want), (metric, far_target, got, want)
    out = U.evaluate(emb, issame)
    want = golden[f"{name}_evaluate_repaired"]                      # mean acc, mean f1, val, std, far
    assert len(out) == 7
    assert np.array_equal(np.array([np.mean(out[2]), np.mean(out[3]), out[4], out[5], out[6]]), want), (out[4:], want)


def test_val_threshold_is_zero_when_the_target_far_is_out_of_reach(monkeypatch, golden):
    """utility.py:108-112: if no threshold of the grid reaches far_target on the train folds, the threshold is 0 and
    nothing is accepted."""
    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.evaluation import utility as U

    emb, issame = pairs(5, 1200, 64, noise=1.0)
    lib = FakeLib({0: golden["hard10_dist0"].astype(np.float32)})
    monkeypatch.setattr(_ffi, "load_library", lambda: lib)
    monkeypatch.setattr(_ffi, "init", lambda device=None: None)
    val, std, far = U.calculate_val(np.arange(0, 0.05, 0.01), emb[0::2], emb[1::2], issame, 0.5)
    assert (val, std, far) == (0.0, 0.0, 0.0)


def test_verify_methods_reproduce_the_reference(monkeypatch, capsys):
    """TripletPrediction.verify / SiamesePrediction.verify against the reference's own methods (predictions.py:104-150,
    :52-89, cut out and executed on a stub self by make_golden_host.py): same distance, same decision, same printed
    line - including the identical-encoding case, where the triplet path reports exactly 0 and the siamese head
    sqrt(K.epsilon()).  dif_euclidean_distance is a numpy stand-in working through the pointers it is handed."""
    import json
    import os

    import torch
    from make_golden_host import verify_cases

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.predictions import SiamesePrediction, TripletPrediction

    class Lib:
        def dif_euclidean_distance(self, x, y, B, D, eps, out, stream):
            a, b = _view(x, (B, D), C.c_float), _view(y, (B, D), C.c_float)
            _view(out, (B,), C.c_float)[:] = np.sqrt(np.maximum(np.sum(np.square(a - b), axis=1), np.float32(eps)))
            return 0

    lib = Lib()
    monkeypatch.setattr(_ffi, "load_library", lambda: lib)
    monkeypatch.setattr(_ffi, "init", lambda device=None: None)
    monkeypatch.setattr(_ffi, "current_stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "host_reference.json")) as f:
        ref = {r["name"]: r for r in json.load(f)["verify"]}
    for name, enc, stored, thr in verify_cases():
        for key, obj, db in (("triplet", TripletPrediction(), {"who": stored[0]}), ("siamese", SiamesePrediction(), {"who": stored})):
            capsys.readouterr()
            dist, ok = obj.verify(enc, "who", db, threshold=thr)
            want = ref[name][key]
            assert capsys.readouterr().out == want["printed"] and ok is want["is_valid"], (name, key)
            assert abs(dist - want["dist"]) <= 1e-6 * max(want["dist"], 1e-3), (name, key, dist, want["dist"])
            if want["dist"] == 0.0:
                assert dist == 0.0
