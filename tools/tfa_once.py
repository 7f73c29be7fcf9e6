"""Time the tfa hard / semi-hard losses (fwd + bwd, device tensors) - development aid.
    python tools/tfa_once.py [P] [K] [D] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deep_insight_face_b200.common.tfa_losses import TFA_HARD, TFA_SEMIHARD, tfa_triplet

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rng = np.random.default_rng(1)
cent = 0.05 * rng.standard_normal((P, D)).astype(np.float32)
x = torch.from_numpy((np.repeat(cent, K, 0) + 0.5 * rng.standard_normal((P * K, D))).astype(np.float32)).cuda()
lab = torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)).cuda()
for kind, name in ((TFA_HARD, "hard"), (TFA_SEMIHARD, "semihard")):
    for _ in range(3):
        tfa_triplet(lab, x, kind)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tfa_triplet(lab, x, kind)
    e1.record()
    torch.cuda.synchronize()
    print("tfa %s B=%d D=%d: %.1f us/step" % (name, P * K, D, e0.elapsed_time(e1) / iters * 1e3))
