"""Drop-in for the loss function of deep_insight_face/networks/triplet.py:16-46 (`triplet_loss`)."""
from __future__ import annotations

import numpy as np

from .. import _ffi


def triplet_loss(y_true, y_pred, alpha=0.4, return_grad=False, dloss=None):
    """Explicit (anchor | positive | negative) triplet loss on concatenated [B, 3D] rows; per-sample [B].
    `y_true` is ignored, as in the reference (networks/triplet.py:16-46)."""
    lib = _ffi.load_library()
    if _ffi.is_device_tensor(y_pred):
        import torch

        y = y_pred.detach().contiguous().float()
        _ffi.init(y.device.index or 0)
        B, D3 = y.shape
        loss = torch.empty(B, dtype=torch.float32, device=y.device)
        dy = torch.empty_like(y) if return_grad else None
        dl = None if dloss is None else dloss.contiguous().float()
        _ffi.check(lib.dif_triplet_apn(_ffi.ptr(y), B, D3 // 3, float(alpha), _ffi.ptr(loss), _ffi.ptr(dl), _ffi.ptr(dy),
                                       _ffi.current_stream_ptr(y.device)))
        return (loss, dy) if return_grad else loss
    import torch

    y = torch.from_numpy(_ffi.host_array(y_pred, np.float32)).cuda()
    out = triplet_loss(None, y, alpha, return_grad, None if dloss is None else torch.from_numpy(
        _ffi.host_array(dloss, np.float32)).cuda())
    if return_grad:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()
