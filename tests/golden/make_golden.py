"""Generates tests/golden/*.npz by running the REFERENCE's own functions (the parts of
deep_insight_face/evaluation/utility.py that import and run in the build container)
on seeded synthetic inputs.  Run once in the build container:

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

calculate_val / evaluate are run with ONE dependency call repaired (scipy's interp1d rejects the repeated x of the FAR
plateau since 1.12, so the shipped function raises at utility.py:109): keys `*_val_repaired*` / `*_evaluate_repaired*`.

/root/reference does not exist on the GPU box; only the .npz files travel.
"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from deep_insight_face.evaluation import utility as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


from synth import pairs  # noqa: E402


def main():
    out = {}
    for name, seed, n_pairs, D in (("small", 4, 600, 128), ("c4", 44, 6000, 128)):
        emb, issame = pairs(seed, n_pairs, D)
        e1, e2 = emb[0::2], emb[1::2]
        out[f"{name}_emb_checksum"] = np.array([emb.astype(np.float64).sum(), float(emb[7, 3]), float(issame.sum())])
        for metric in (0, 1):
            d = ref.distance(e1, e2, metric)
            out[f"{name}_dist{metric}"] = d
            thr = np.arange(0, 4, 0.01)
            acc = np.array([ref.calculate_accuracy(t, d, issame) for t in thr[:: 25]])
            out[f"{name}_acc{metric}"] = acc
            vf = np.array([ref.calculate_val_far(t, d, issame) for t in thr[:: 25]])
            out[f"{name}_valfar{metric}"] = vf
            for sm in (False, True):
                with contextlib.redirect_stdout(io.StringIO()):
                    tpr, fpr, accuracy, f1 = ref.calculate_roc(thr, e1, e2, issame, nrof_folds=10,
                                                               distance_metric=metric, subtract_mean=sm)
                out[f"{name}_roc{metric}_{int(sm)}_tpr"] = tpr
                out[f"{name}_roc{metric}_{int(sm)}_fpr"] = fpr
                out[f"{name}_roc{metric}_{int(sm)}_acc"] = accuracy
                out[f"{name}_roc{metric}_{int(sm)}_f1"] = f1
    # calculate_val / evaluate (utility.py:80-119, :10-33).  As shipped the function stops at line 109 on scipy >= 1.12:
    # interp1d rejects the repeated x values every far_train has (its plateau at 0).  ONE dependency call is repaired
    # for the run - interp1d(x, y, kind='slinear') is replaced by the same piecewise-linear inverse with repeated x
    # allowed (np.interp) - everything else is the reference's own code.  The keys say so: `*_val_repaired*`.
    class _Interp:
        @staticmethod
        def interp1d(x, y, kind="slinear"):
            assert kind == "slinear"
            return lambda t: np.interp(t, x, y)

    saved = ref.interpolate
    ref.interpolate = _Interp
    try:
        for name, seed, n_pairs, D in (("small", 4, 600, 128), ("c4", 44, 6000, 128)):
            emb, issame = pairs(seed, n_pairs, D)
            e1, e2 = emb[0::2], emb[1::2]
            thr = np.arange(0, 4, 0.001)
            for metric in (0, 1):
                for sm in (False, True):
                    out[f"{name}_val_repaired{metric}_{int(sm)}"] = np.array(
                        ref.calculate_val(thr, e1, e2, np.asarray(issame), 1e-3, nrof_folds=10, distance_metric=metric,
                                          subtract_mean=sm))
            out[f"{name}_val_repaired_far1e-2"] = np.array(ref.calculate_val(thr, e1, e2, np.asarray(issame), 1e-2))
            with contextlib.redirect_stdout(io.StringIO()):
                ev = ref.evaluate(emb, issame)
            out[f"{name}_evaluate_repaired_val"] = np.array(ev[4:7])
        # the two sets above separate perfectly (VAL = 1); overlapping classes make VAL and its spread informative
        for name, seed, n_pairs, D, noise in (("hard10", 5, 1200, 64, 1.0), ("hard15", 6, 1200, 64, 1.5)):
            emb, issame = pairs(seed, n_pairs, D, noise=noise)
            e1, e2 = emb[0::2], emb[1::2]
            thr = np.arange(0, 4, 0.001)
            for metric in (0, 1):
                for sm in (False, True):
                    for far_target in (1e-3, 1e-2):
                        out[f"{name}_val_repaired{metric}_{int(sm)}_far{far_target:g}"] = np.array(
                            ref.calculate_val(thr, e1, e2, np.asarray(issame), far_target, nrof_folds=10,
                                              distance_metric=metric, subtract_mean=sm))
                out[f"{name}_dist{metric}"] = ref.distance(e1, e2, metric)
            with contextlib.redirect_stdout(io.StringIO()):
                ev = ref.evaluate(emb, issame)
            out[f"{name}_evaluate_repaired"] = np.concatenate([[np.mean(ev[2]), np.mean(ev[3])], ev[4:7]])   # acc, f1, val, std, far
    finally:
        ref.interpolate = saved
    np.savez_compressed(os.path.join(HERE, "verification_reference.npz"), **out)
    print("wrote", os.path.join(HERE, "verification_reference.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
