"""Generates tests/golden/*.npz by running the REFERENCE's own functions (the parts of
deep_insight_face/evaluation/utility.py that import and run in the build container)
on seeded synthetic inputs.  Run once in the build container:

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

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
    np.savez_compressed(os.path.join(HERE, "verification_reference.npz"), **out)
    print("wrote", os.path.join(HERE, "verification_reference.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
