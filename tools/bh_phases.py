"""Development aid: in-kernel phase stamps of the one-launch batch-hard step (DIF_BH_PROFILE=1)."""
import os
import sys

os.environ["DIF_BH_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deep_insight_face_b200 import _ffi
from deep_insight_face_b200.common.losses import BatchHardStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 72
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
variant = int(sys.argv[3]) if len(sys.argv) > 3 else _ffi.LOSS_BH_COSINE
rng = np.random.default_rng(1)
P, K = B // 4, 4
emb = (np.repeat(rng.standard_normal((P, D)), K, 0) + rng.standard_normal((B, D))).astype(np.float32)
step = BatchHardStep(B, D, variant, 0.35, "cuda:0", graph=False)
step.emb.copy_(torch.from_numpy(emb))
step.labels.copy_(torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)))
for _ in range(6):
    step()
torch.cuda.synchronize()
