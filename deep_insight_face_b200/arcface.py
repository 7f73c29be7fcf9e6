"""ArcFace additive-angular-margin logits + softmax cross-entropy on B200 (csrc/arcface.cu).

The reference has no classification head beyond `l2_normalize` (networks/triplet.py:138, inceptionv3.py:305);
this is the margin head named by BASELINE.json, specified in DESIGN.md (arXiv 1801.07698):
logits = s * cos(theta + m * onehot) over L2-normalised embeddings X [B, D] and class centres W [C, D].
"""
from __future__ import annotations

import numpy as np

from . import _ffi


def arcface_loss(X, W, y, s: float = 64.0, m: float = 0.5, dloss=None, want_grad: bool = True,
                 precision="tf32x3"):
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
                                   _ffi.ptr(dl), _ffi.ptr(dX), _ffi.ptr(dW), _ffi.precision_code(precision),
                                   _ffi.current_stream_ptr(x.device)))
        return (loss, dX, dW) if want_grad else loss
    _ffi.init()
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
                                    _ffi.ptr(dl), _ffi.ptr(dX), _ffi.ptr(dW), _ffi.precision_code(precision)))
    return (loss, dX, dW) if want_grad else loss


class ArcFaceStep:
    """Preallocated, optionally CUDA-graphed ArcFace fwd + bwd for fixed (B, C, D): the training-loop form.

    One step is 13 kernel launches and a dozen tensor-map encodes; with `graph=True` they are captured once and
    replayed with a single cudaGraphLaunch.  Write into `X`, `W`, `y` in place, call the step, read `loss`, `dX`, `dW`.
    """

    def __init__(self, B: int, C: int, D: int, s: float = 64.0, m: float = 0.5, device="cuda:0", graph: bool = False,
                 precision="tf32x3"):
        self.precision = _ffi.precision_code(precision)
        import torch

        dev = torch.device(device)
        _ffi.init(dev.index or 0)
        self._lib = _ffi.load_library()
        self.B, self.C, self.D, self.s, self.m = int(B), int(C), int(D), float(s), float(m)
        self.X = torch.zeros((B, D), dtype=torch.float32, device=dev)
        self.W = torch.zeros((C, D), dtype=torch.float32, device=dev)
        self.y = torch.zeros(B, dtype=torch.int32, device=dev)
        self.loss = torch.empty(B, dtype=torch.float32, device=dev)
        self.dX = torch.empty_like(self.X)
        self.dW = torch.empty_like(self.W)
        self._dev = dev
        self._graph = None
        self.X.normal_()
        self.W.normal_()
        self._launch()                      # warm-up: sizes the library workspace outside any capture
        torch.cuda.synchronize(dev)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch()
            self._graph = g

    def _launch(self):
        import torch

        st = int(torch.cuda.current_stream(self._dev).cuda_stream)
        _ffi.check(self._lib.dif_arcface(self.X.data_ptr(), self.W.data_ptr(), self.y.data_ptr(), self.B, self.C, self.D,
                                         self.s, self.m, self.loss.data_ptr(), None, self.dX.data_ptr(),
                                         self.dW.data_ptr(), self.precision, st))

    def __call__(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._launch()
        return self.loss, self.dX, self.dW


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
