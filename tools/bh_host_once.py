"""Host-to-host time of the C1 batch-hard step (numpy in, numpy out) - development aid.
python tools/bh_host_once.py [B] [D]   (DIF_BH_HOST_STAGED=1 selects the copy-in / copy-out path)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from deep_insight_face_b200 import _ffi
from deep_insight_face_b200.common.losses import BatchHardTripletLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 72
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(1)
P, K = B // 4, 4
emb = (np.repeat(rng.standard_normal((P, D)), K, 0) + rng.standard_normal((B, D))).astype(np.float32)
lab = np.repeat(np.arange(P), K).astype(np.int32)
loss = BatchHardTripletLoss()
for _ in range(20):
    ref = loss.loss_and_grad(lab, emb)
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    loss.loss_and_grad(lab, emb)
t_api = (time.perf_counter() - t0) / n * 1e6
# the same call with the outputs and the argument tuple built once
lib = _ffi.load_library()
lo, po, ne = np.empty(B, np.float32), np.empty(B, np.int32), np.empty(B, np.int32)
st, gr = np.empty(4, np.float32), np.empty((B, D), np.float32)
args = (_ffi.ptr(emb), _ffi.ptr(lab), B, D, _ffi.LOSS_BH_COSINE, 0.35, _ffi.ptr(lo), _ffi.ptr(po), _ffi.ptr(ne), _ffi.ptr(st),
        None, _ffi.ptr(gr), _ffi.PREC_TF32X3)
fn = lib.dif_batch_hard_host
for _ in range(20):
    fn(*args)
t0 = time.perf_counter()
for _ in range(n):
    fn(*args)
t_raw = (time.perf_counter() - t0) / n * 1e6
from deep_insight_face_b200.common.losses import BatchHardHostStep

hs = BatchHardHostStep(B, D, _ffi.LOSS_BH_COSINE, 0.35)
hs.emb[:] = emb
hs.labels[:] = lab
for _ in range(20):
    hs()
t0 = time.perf_counter()
for _ in range(n):
    hs()
t_step = (time.perf_counter() - t0) / n * 1e6
print("  BatchHardHostStep (works in the page-locked block, no copies): %.1f us, equal: %s" % (t_step, np.array_equal(hs.grad, ref[1])))
ok = np.array_equal(lo, ref[0]) and np.array_equal(gr, ref[1]) and np.array_equal(po, ref[2]["pos_idx"])
print("batch-hard host-to-host B=%d D=%d: %.1f us through loss_and_grad, %.1f us raw C call, outputs equal: %s, staged=%s"
      % (B, D, t_api, t_raw, ok, bool(os.environ.get("DIF_BH_HOST_STAGED"))))
