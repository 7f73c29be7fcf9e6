"""ArcFace additive-angular-margin logits + softmax cross-entropy on B200 (csrc/arcface.cu).

The reference has no classification head beyond `l2_normalize` (networks/triplet.py:138, inceptionv3.py:305);
this is the margin head named by BASELINE.json, specified in DESIGN.md (arXiv 1801.07698):
logits = s * cos(theta + m * onehot) over L2-normalised embeddings X [B, D] and class centres W [C, D].
"""
from __future__ import annotations

import numpy as np

from . import _ffi


def arcface_loss(X, W, y, s: float = 64.0, m: float = 0.5, dloss=None, want_grad: bool = True):
    """Per-sample loss [B] and (optionally) dX [B, D], dW [C, D] = gradients of mean(loss) (or of sum dloss * loss).

    numpy / torch-CPU inputs take the host entry point (pinned staging, H2D, D2H); torch-CUDA inputs stay on
    the device and run on torch's current stream.
    """
    lib = _ffi.load_library()
    if _ffi.is_device_tensor(X):
        import torch

        x = X.detach().contiguous().float()
        w = W.detach().contiguous().float()
        yy = y.detach().to(torch.int32).contiguous()
        _ffi.init(x.device.index or 0)
        B, D = x.shape
        C = w.shape[0]
        loss = torch.empty(B, dtype=torch.float32, device=x.device)
        dX = torch.empty_like(x) if want_grad else None
        dW = torch.empty_like(w) if want_grad else None
        dl = None if dloss is None else dloss.detach().contiguous().float()
        _ffi.check(lib.dif_arcface(_ffi.ptr(x), _ffi.ptr(w), _ffi.ptr(yy), B, C, D, float(s), float(m), _ffi.ptr(loss),
                                   _ffi.ptr(dl), _ffi.ptr(dX), _ffi.ptr(dW), _ffi.PREC_TF32X3,
                                   _ffi.current_stream_ptr(x.device)))
        return (loss, dX, dW) if want_grad else loss
    _ffi.init(0)
    x = _ffi.host_array(X, np.float32)
    w = _ffi.host_array(W, np.float32)
    yy = np.ascontiguousarray(_ffi.host_array(y, None), dtype=np.int32)
    B, D = x.shape
    C = w.shape[0]
    if w.shape[1] != D or yy.shape != (B,):
        raise ValueError("X [B, D], W [C, D], y [B] expected")
    if yy.min() < 0 or yy.max() >= C:
        raise ValueError("labels out of range")
    loss = np.empty(B, dtype=np.float32)
    dX = np.empty_like(x) if want_grad else None
    dW = np.empty_like(w) if want_grad else None
    dl = None if dloss is None else _ffi.host_array(dloss, np.float32, (B,))
    _ffi.check(lib.dif_arcface_host(_ffi.ptr(x), _ffi.ptr(w), _ffi.ptr(yy), B, C, D, float(s), float(m), _ffi.ptr(loss),
                                    _ffi.ptr(dl), _ffi.ptr(dX), _ffi.ptr(dW), _ffi.PREC_TF32X3))
    return (loss, dX, dW) if want_grad else loss


class ArcFaceLoss:
    """Keras-style callable: `loss(y_true, y_pred)` with y_pred = embeddings and the class centres held here."""

    def __init__(self, weights, s: float = 64.0, m: float = 0.5):
        self.weights = weights
        self.s, self.m = float(s), float(m)

    def call(self, y_true, embeddings):
        y = _ffi.host_array(y_true, None) if not _ffi.is_device_tensor(y_true) else y_true
        if getattr(y, "ndim", 1) == 2 or (hasattr(y, "dim") and y.dim() == 2):
            y = y.argmax(1)
        return arcface_loss(embeddings, self.weights, y, self.s, self.m, want_grad=False)

    def __call__(self, y_true, embeddings):
        return self.call(y_true, embeddings).mean()

    def get_config(self):
        return {"s": self.s, "m": self.m}
