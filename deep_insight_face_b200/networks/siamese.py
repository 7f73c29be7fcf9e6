"""Drop-in for the distance / loss functions of deep_insight_face/networks/siamese.py:22-45."""
from __future__ import annotations

import numpy as np

from .. import _ffi

EPSILON = 1e-7  # K.epsilon()


def _cuda(a):
    import torch

    if _ffi.is_device_tensor(a):
        return a.contiguous().float(), True
    return torch.from_numpy(_ffi.host_array(a, np.float32)).cuda(), False


def euclidean_distance(vects):
    """siamese.py:22-24: sqrt(max(sum((x - y)^2, axis=1, keepdims=True), K.epsilon())) -> [B, 1]."""
    import torch

    x, on_dev = _cuda(vects[0])
    y, _ = _cuda(vects[1])
    lib = _ffi.load_library()
    _ffi.init(x.device.index or 0)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _ffi.check(lib.dif_euclidean_distance(_ffi.ptr(x), _ffi.ptr(y), x.shape[0], x.shape[1], EPSILON, _ffi.ptr(out),
                                          _ffi.current_stream_ptr(x.device)))
    out = out[:, None]
    return out if on_dev else out.cpu().numpy()


def contrastive_loss(y_true, y_pred, return_grad=False):
    """siamese.py:32-39: mean(y * d^2 + (1 - y) * max(1 - d, 0)^2)."""
    import torch

    yt, on_dev = _cuda(np.reshape(y_true, -1) if not _ffi.is_device_tensor(y_true) else y_true.reshape(-1))
    d, _ = _cuda(np.reshape(y_pred, -1) if not _ffi.is_device_tensor(y_pred) else y_pred.reshape(-1))
    lib = _ffi.load_library()
    _ffi.init(d.device.index or 0)
    out = torch.empty(1, dtype=torch.float32, device=d.device)
    dd = torch.empty_like(d) if return_grad else None
    _ffi.check(lib.dif_contrastive_loss(_ffi.ptr(yt), _ffi.ptr(d), d.shape[0], 1.0, _ffi.ptr(out), _ffi.ptr(dd),
                                        _ffi.current_stream_ptr(d.device)))
    val = out[0] if on_dev else float(out.item())
    if return_grad:
        return val, (dd if on_dev else dd.cpu().numpy())
    return val


def _accuracy(y_true, y_pred, threshold=0.4):
    """siamese.py:42-45: mean(y_true == (y_pred < threshold)), default threshold 0.4 as in the reference (it is
    compiled as `metrics=[_accuracy]`, siamese.py:158, so the default is what training reports); host-side."""
    y = np.reshape(_ffi.host_array(y_true, None), -1)
    d = np.reshape(_ffi.host_array(y_pred, None), -1)
    return float(np.mean(y == (d < threshold).astype(y.dtype)))
