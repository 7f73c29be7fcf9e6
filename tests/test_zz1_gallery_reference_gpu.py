"""The CUDA 1:N search against rankings made by the REFERENCE's own distance function
(evaluation/utility.py:52-66 run over every (query, gallery row) pair by tests/golden/make_golden_gallery.py):
same rows in the same order wherever the reference's distances decide the order, same numbers (squared-L2 scores are
the metric-0 distances, cosine scores are cos(pi * metric-1 distance)).  The bit-exact comparison with the oracle
is tests/test_gallery_gpu.py; this file ties the same call to the reference itself."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("precision", ["tf32x3", "bf16"])
def test_search_ranks_like_the_reference_distance(gpu, precision):
    from synth import GALLERY_CASES, check_ranking, gallery_case

    from deep_insight_face_b200.gallery import Gallery

    ref = np.load(os.path.join(HERE, "golden", "gallery_reference.npz"))
    for name, seed, N, Q, D, k in GALLERY_CASES:
        rows, q, pick = gallery_case(seed, N, Q, D)
        for metric, mname in ((0, "l2"), (1, "cosine")):
            with Gallery(N, D, mname, precision) as g:
                g.add(rows)
                s, _, r = g.search(q, k, return_rows=True)
            share = check_ranking(ref[f"{name}/metric{metric}/dist"], ref[f"{name}/metric{metric}/rows"],
                                  np.asarray(s), np.asarray(r).astype(np.int64), metric)
            assert share > 0.97, (name, metric, share)
            assert (np.asarray(r)[:, 0] == pick).all()
