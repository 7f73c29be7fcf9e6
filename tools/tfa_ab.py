"""Development aid: run the tfa losses over a few shapes and save every output, to compare two builds / switches bit for bit.
    python tools/tfa_ab.py out.npz ;  DIF_TFA_SLOW_PDIST=1 python tools/tfa_ab.py ref.npz ;  python tools/tfa_ab.py --cmp out.npz ref.npz"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not a[k].tobytes() == b[k].tobytes()]
    print("compared %d arrays: %s" % (len(a.files), "ALL BIT-IDENTICAL" if not bad else "DIFFERENT: %s" % bad))
    sys.exit(1 if bad else 0)
import torch

from deep_insight_face_b200.common.tfa_losses import TFA_HARD, TFA_SEMIHARD, tfa_triplet

out = {}
rng = np.random.default_rng(7)
for (P, K, D) in ((18, 4, 128), (83, 4, 128), (256, 4, 128), (100, 3, 64), (50, 5, 100), (40, 2, 32), (33, 3, 20), (512, 4, 96), (150, 2, 256), (70, 4, 512), (90, 3, 200), (300, 1, 300)):
    B = P * K
    cent = 0.05 * rng.standard_normal((P, D)).astype(np.float32)
    x = (np.repeat(cent, K, 0) + 0.5 * rng.standard_normal((B, D))).astype(np.float32)
    x[5] = x[4]          # exact duplicates: zero distances, ties
    x[B - 1] = x[0]
    xt = torch.from_numpy(x).cuda()
    lab = torch.from_numpy(np.repeat(np.arange(P), K).astype(np.int32)).cuda()
    for kind, name in ((TFA_HARD, "hard"), (TFA_SEMIHARD, "semi")):
        res = tfa_triplet(lab, xt, kind)
        flat = []
        def walk(o, key):
            if isinstance(o, dict):
                for k, v in o.items():
                    walk(v, key + "_" + str(k))
            elif isinstance(o, (tuple, list)):
                for i, v in enumerate(o):
                    walk(v, key + "_" + str(i))
            elif hasattr(o, "detach"):
                out[key] = o.detach().cpu().numpy()
            elif o is not None:
                out[key] = np.asarray(o)
        walk(res, "%s_P%d_K%d_D%d" % (name, P, K, D))
np.savez(sys.argv[1], **out)
print("saved %d arrays" % len(out))
