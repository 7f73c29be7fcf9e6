"""Runs the loss kernels a few times (for `ncu` launch lists): loss_once.py [bh72|bh4096|arcface|ball4096]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deep_insight_face_b200 import _ffi
from deep_insight_face_b200.arcface import arcface_loss
from deep_insight_face_b200.common.losses import batch_hard, batch_all
what = sys.argv[1] if len(sys.argv) > 1 else "bh4096"
rng = np.random.default_rng(1)
if what.startswith("bh") or what.startswith("ball"):
    B = int(what.lstrip("bhal"))
    P, K, D = B // 4, 4, 128
    emb = torch.from_numpy((np.repeat(rng.standard_normal((P, D)), K, 0) + rng.standard_normal((B, D))).astype(np.float32)).cuda()
    lab = torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)).cuda()
    for _ in range(3):
        if what.startswith("ball"):
            batch_all(lab, emb, 0.35)
        else:
            batch_hard(lab, emb, _ffi.LOSS_BH_COSINE, 0.35)
else:
    B, C, D = 512, 10000, 512
    X = torch.randn(B, D, device="cuda"); W = 0.01 * torch.randn(C, D, device="cuda"); y = torch.randint(0, C, (B,), device="cuda")
    for _ in range(3):
        arcface_loss(X, W, y)
torch.cuda.synchronize()
print("ok")
