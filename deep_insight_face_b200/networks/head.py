"""Embedding head: the `l2_normalize` Lambda that ends the reference's embedding models
(deep_insight_face/networks/inceptionv3.py:305 `K.l2_normalize(x, 1)`, networks/triplet.py:138
`tf.compat.v1.nn.l2_normalize(x, axis=1)`), in the canonical fp32 arithmetic the gallery and the losses use, so
an embedding normalised here is bit-identical to the row `Gallery.add` stores for it."""
from __future__ import annotations

import numpy as np

from .. import _ffi


def _launch_fwd(x):
    import torch

    lib = _ffi.load_library()
    _ffi.init(x.device.index or 0)
    y = torch.empty_like(x)
    inv = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _ffi.check(lib.dif_l2_normalize(_ffi.ptr(x), x.shape[0], x.shape[1], _ffi.ptr(y), _ffi.ptr(inv),
                                    _ffi.current_stream_ptr(x.device)))
    return y, inv


def _launch_bwd(g, y, inv):
    """d / dx of y = l2_normalize(x) applied to the cotangent g [n, D] (device tensors; y, inv from _launch_fwd)."""
    import torch

    g = g.contiguous().float()
    dx = torch.empty_like(y)
    _ffi.check(_ffi.load_library().dif_l2_normalize_bwd(_ffi.ptr(g), _ffi.ptr(y), _ffi.ptr(inv), y.shape[0], y.shape[1],
                                                        _ffi.ptr(dx), _ffi.current_stream_ptr(y.device)))
    return dx


def l2_normalize(x, axis=1):
    """x [n, D] -> x * rsqrt(max(sum(x^2, axis=1), 1e-12)).  numpy in -> numpy out; torch-CUDA in -> torch out,
    differentiable (the backward is dif_l2_normalize_bwd)."""
    if axis not in (1, -1):
        raise ValueError("only row-wise normalisation (axis=1) is on the GPU path, as the reference uses it")
    import torch

    if not _ffi.is_device_tensor(x):
        arr = _ffi.host_array(x, np.float32)
        if arr.ndim != 2:
            raise ValueError("x must be [n, D]")
        return _launch_fwd(torch.from_numpy(arr).cuda())[0].cpu().numpy()
    if x.dim() != 2:
        raise ValueError("x must be [n, D]")

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp):
            y, inv = _launch_fwd(inp.detach().contiguous().float())
            ctx.save_for_backward(y, inv)
            return y

        @staticmethod
        def backward(ctx, g):
            y, inv = ctx.saved_tensors
            return _launch_bwd(g, y, inv)

    return _Fn.apply(x)
