"""Time the CUDA-graphed batch-hard step (fwd + bwd) - development aid.   python tools/bh_once.py [B] [D] [variant] [path]
(path: 0 auto, 1 CUDA-core miner, 2 tensor-core miner, 3 one-launch cluster step - dif_batch_hard_set_path)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deep_insight_face_b200 import _ffi
from deep_insight_face_b200.common.losses import BatchHardStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
variant = int(sys.argv[3]) if len(sys.argv) > 3 else _ffi.LOSS_BH_COSINE
if len(sys.argv) > 4:
    _ffi.init(0)
    _ffi.check(_ffi.load_library().dif_batch_hard_set_path(int(sys.argv[4])))
rng = np.random.default_rng(1)
P, K = B // 4, 4
emb = (np.repeat(rng.standard_normal((P, D)), K, 0) + rng.standard_normal((B, D))).astype(np.float32)
step = BatchHardStep(B, D, variant, 0.35, "cuda:0", graph=True)
step.emb.copy_(torch.from_numpy(emb))
step.labels.copy_(torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)))
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    step()
e1.record()
torch.cuda.synchronize()
print("batch-hard variant %d B=%d D=%d: %.1f us/step" % (variant, B, D, e0.elapsed_time(e1) * 10))
