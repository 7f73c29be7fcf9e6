"""Generates tests/golden/gallery_reference.npz: 1:N rankings by the REFERENCE's own distance function.

    python tests/golden/make_golden_gallery.py          (in the build container)

The reference has no 1:N search (oneshot.py:8 is a stub); what it has is the distance every
comparison goes through, evaluation/utility.py:52-66 `distance(embeddings1, embeddings2, distance_metric)`:
metric 0 = sum of squared differences, metric 1 = arccos(cosine similarity) / pi.  A 1:N search by that function is
"probe the query against every gallery row with `distance`, keep the k smallest"; this script does exactly that
with the imported, unmodified function (the query broadcast against the gallery, float64 copies of the fp32 inputs
so the reference's formula is evaluated without fp32 summation noise) and stores, per query, the k + 1 smallest
distances and their rows (stable order: ties to the lower row).  tests/test_parity_cpu.py holds
oracle.c_oracle.gallery_search to it, tests/test_gallery_gpu.py the CUDA path.  Only the .npz travels to the GPU box.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from deep_insight_face.evaluation import utility as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import GALLERY_CASES, gallery_case  # noqa: E402

def main():
    out = {}
    for name, seed, N, Q, D, k in GALLERY_CASES:
        rows, q, pick = gallery_case(seed, N, Q, D)
        g64 = rows.astype(np.float64)
        for metric in (0, 1):
            dist = np.empty((Q, k + 1))
            idx = np.empty((Q, k + 1), dtype=np.int64)
            for i in range(Q):
                d = ref.distance(np.broadcast_to(q[i].astype(np.float64), g64.shape), g64, metric)  # the reference
                order = np.argsort(d, kind="stable")[: k + 1]
                dist[i], idx[i] = d[order], order
            out[f"{name}/metric{metric}/dist"] = dist
            out[f"{name}/metric{metric}/rows"] = idx
        out[f"{name}/checksum"] = np.array([rows.astype(np.float64).sum(), q.astype(np.float64).sum(), float(pick.sum())])
    path = os.path.join(HERE, "gallery_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
