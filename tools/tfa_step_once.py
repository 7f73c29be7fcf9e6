"""Time the CUDA-graphed tfa step (TfaTripletStep) - development aid.   python tools/tfa_step_once.py [P] [K] [D]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deep_insight_face_b200.common.tfa_losses import TFA_HARD, TFA_SEMIHARD, TfaTripletStep

P = int(sys.argv[1]) if len(sys.argv) > 1 else 18
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rng = np.random.default_rng(1)
B = P * K
x = torch.from_numpy((np.repeat(0.05 * rng.standard_normal((P, D)), K, 0) + 0.5 * rng.standard_normal((B, D))).astype(np.float32)).cuda()
lab = torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)).cuda()
for kind, name in ((TFA_HARD, "hard"), (TFA_SEMIHARD, "semihard")):
    step = TfaTripletStep(B, D, kind, 1.0, "cuda:0", graph=True)
    step.emb.copy_(x)
    step.labels.copy_(lab)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("tfa %s graphed B=%d D=%d: %.1f us/step" % (name, B, D, e0.elapsed_time(e1) * 5))
