"""Small-shape tour of every kernel family for `compute-sanitizer --tool memcheck`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
from deep_insight_face_b200 import _ffi
from deep_insight_face_b200.gallery import Gallery
from deep_insight_face_b200.common.losses import (BatchHardTripletLoss, BatchHardTripletLossEuclidean, BatchAllTripletLoss)
from deep_insight_face_b200.arcface import arcface_loss
from deep_insight_face_b200.evaluation import utility as U
from deep_insight_face_b200.networks.triplet import triplet_loss
from synth import pairs

_ffi.init(0)
lib = _ffi.load_library()
rng = np.random.default_rng(0)
for prec in ("tf32x3", "bf16", "tf32x1"):
    for metric in ("cosine", "l2"):
        rows = rng.standard_normal((1337, 96)).astype(np.float32)
        q = rows[:37] + 0.1 * rng.standard_normal((37, 96)).astype(np.float32)
        with Gallery(2000, 96, metric, prec) as g:
            g.add(rows)
            s, ids = g.search(q, 7)
            assert (ids[:, 0] == np.arange(37)).all()
            g.set_option("force_fallback", 1)
            g.search(q, 7)
emb = rng.standard_normal((100, 64)).astype(np.float32)
lab = np.repeat(np.arange(25), 4)
for path in (1, 2):
    _ffi.check(lib.dif_batch_hard_set_path(path))
    BatchHardTripletLoss().loss_and_grad(lab, emb)
    BatchHardTripletLossEuclidean(alpha=30.0).loss_and_grad(lab, emb)
_ffi.check(lib.dif_batch_hard_set_path(0))
BatchAllTripletLoss().loss_and_grad(lab, emb)
X = rng.standard_normal((70, 64)).astype(np.float32); W = (0.01 * rng.standard_normal((301, 64))).astype(np.float32)
arcface_loss(X, W, rng.integers(0, 301, 70))
e, issame = pairs(4, 600, 128)
U.evaluate(e, issame)
triplet_loss(None, rng.standard_normal((9, 3 * 40)).astype(np.float32), return_grad=True)
print("tour ok")
