"""Drop-in for deep_insight_face/common/losses.py: the batch-hard triplet losses on B200 kernels.

Same class names, constructor arguments, `call(labels, embeddings)` contract (one-hot labels, per-sample
[B] result, reduction left to Keras) and `get_config` / `from_config` round trip as the reference
(common/losses.py:5-30).  The arithmetic runs in libdif_b200.so (csrc/batch_hard.cu); there is no CPU path.

    call(labels, embeddings)            per-sample loss; numpy in -> numpy out, torch-CUDA in -> differentiable
                                        torch tensor out, TensorFlow in (when TF is installed) -> tf.custom_gradient
    loss_and_grad(labels, embeddings)   (loss [B], d mean(loss) / d embeddings [B, D], info) in one fused call
"""
from __future__ import annotations

import numpy as np

from .. import _ffi

try:  # the reference subclasses tf.keras.losses.Loss; TensorFlow is optional here
    import tensorflow as _tf  # type: ignore

    _LossBase = _tf.keras.losses.Loss
except Exception:  # pragma: no cover - TensorFlow is not installed in the build container
    _tf = None

    class _LossBase:  # the slice of the Keras Loss protocol the reference uses
        def __init__(self, reduction="auto", name=None, **kwargs):
            self.reduction = reduction
            self.name = name or type(self).__name__

        def __call__(self, y_true, y_pred, sample_weight=None):
            per_sample = self.call(y_true, y_pred)
            if sample_weight is not None:
                per_sample = per_sample * sample_weight
            return per_sample.mean()  # Keras AUTO / SUM_OVER_BATCH_SIZE

        def get_config(self):
            return {"reduction": self.reduction, "name": self.name}


def _int_labels(labels):
    """One-hot [B, C] (common/losses.py:35 argmax's it) or sparse int [B] labels -> host int32 [B]."""
    lab = _ffi.host_array(labels, None)
    if lab.ndim == 2:
        lab = np.argmax(lab, axis=1)
    return np.ascontiguousarray(lab, dtype=np.int32)


def batch_hard(labels, embeddings, variant: int, alpha: float, dloss=None, want_grad: bool = True):
    """Framework-neutral core.  Host arrays take dif_batch_hard_host (pinned staging, H2D, D2H); torch-CUDA
    embeddings stay on the device.  Returns (loss, grad or None, info)."""
    lib = _ffi.load_library()
    if _ffi.is_device_tensor(embeddings):
        import torch

        emb = embeddings.detach().contiguous().float()
        dev = emb.device
        _ffi.init(dev.index or 0)
        B, D = emb.shape
        st = _ffi.current_stream_ptr(dev)
        if _ffi.is_device_tensor(labels) and labels.dim() == 2:
            oh = labels.detach().contiguous().float()
            lab = torch.empty(B, dtype=torch.int32, device=dev)
            _ffi.check(lib.dif_labels_from_onehot(_ffi.ptr(oh), B, oh.shape[1], _ffi.ptr(lab), st))
        elif _ffi.is_device_tensor(labels):
            lab = labels.detach().to(torch.int32).contiguous()
        else:
            lab = torch.from_numpy(_int_labels(labels)).to(dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        pos = torch.empty(B, dtype=torch.int32, device=dev)
        neg = torch.empty(B, dtype=torch.int32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        grad = torch.empty_like(emb) if want_grad else None
        dl = None if dloss is None else dloss.detach().contiguous().float()
        _ffi.check(lib.dif_batch_hard(_ffi.ptr(emb), _ffi.ptr(lab), B, D, variant, float(alpha), _ffi.ptr(loss),
                                      _ffi.ptr(pos), _ffi.ptr(neg), _ffi.ptr(stats), _ffi.ptr(dl), _ffi.ptr(grad),
                                      _ffi.PREC_TF32X3, st))
        return loss, grad, {"pos_idx": pos, "neg_idx": neg, "stats": stats}
    _ffi.init()
    emb = _ffi.host_array(embeddings, np.float32)
    if emb.ndim != 2:
        raise ValueError("embeddings must be [B, D]")
    B, D = emb.shape
    lab = _int_labels(labels)
    if lab.shape != (B,):
        raise ValueError(f"labels must be one-hot [B, C] or int [B]; got {lab.shape} for B={B}")
    loss = np.empty(B, dtype=np.float32)
    pos = np.empty(B, dtype=np.int32)
    neg = np.empty(B, dtype=np.int32)
    stats = np.empty(4, dtype=np.float32)
    grad = np.empty((B, D), dtype=np.float32) if want_grad else None
    dl = None if dloss is None else _ffi.host_array(dloss, np.float32, (B,))
    _ffi.check(lib.dif_batch_hard_host(_ffi.ptr(emb), _ffi.ptr(lab), B, D, variant, float(alpha), _ffi.ptr(loss),
                                       _ffi.ptr(pos), _ffi.ptr(neg), _ffi.ptr(stats), _ffi.ptr(dl), _ffi.ptr(grad),
                                       _ffi.PREC_TF32X3))
    return loss, grad, {"pos_idx": pos, "neg_idx": neg, "stats": stats}


class BatchHardStep:
    """Preallocated, optionally CUDA-graphed batch-hard step for a fixed (B, D): the training-loop form.

    The three kernels of a step take ~12 us on a B200 at B = 72; building five output tensors and a 14-argument
    ctypes call per step costs more than that, so the buffers and the argument tuple are made once and a step is
    one foreign call - or, with `graph=True`, one cudaGraphLaunch of the captured kernel sequence.

        step = BatchHardStep(B, D, variant=LOSS_BH_COSINE, alpha=0.35, device="cuda:0", graph=True)
        step.emb.copy_(embeddings); step.labels.copy_(int_labels)      # or write into them in place
        step()                                                          # fills step.loss, step.grad, step.pos_idx, ...
    """

    def __init__(self, B: int, D: int, variant: int = _ffi.LOSS_BH_COSINE, alpha: float = 0.35, device="cuda:0",
                 graph: bool = False, want_grad: bool = True):
        import torch

        dev = torch.device(device)
        _ffi.init(dev.index or 0)
        self._lib = _ffi.load_library()
        self.B, self.D, self.variant, self.alpha = int(B), int(D), int(variant), float(alpha)
        self.emb = torch.zeros((B, D), dtype=torch.float32, device=dev)
        self.labels = torch.zeros(B, dtype=torch.int32, device=dev)
        self.loss = torch.empty(B, dtype=torch.float32, device=dev)
        self.pos_idx = torch.empty(B, dtype=torch.int32, device=dev)
        self.neg_idx = torch.empty(B, dtype=torch.int32, device=dev)
        self.stats = torch.empty(4, dtype=torch.float32, device=dev)
        self.grad = torch.empty((B, D), dtype=torch.float32, device=dev) if want_grad else None
        self._dev = dev
        self._graph = None
        self._launch()                       # warm-up: sizes the library workspace outside any capture
        torch.cuda.synchronize(dev)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch()
            self._graph = g

    def _launch(self):
        import torch

        st = int(torch.cuda.current_stream(self._dev).cuda_stream)
        _ffi.check(self._lib.dif_batch_hard(self.emb.data_ptr(), self.labels.data_ptr(), self.B, self.D, self.variant,
                                            self.alpha, self.loss.data_ptr(), self.pos_idx.data_ptr(),
                                            self.neg_idx.data_ptr(), self.stats.data_ptr(), None,
                                            None if self.grad is None else self.grad.data_ptr(), _ffi.PREC_TF32X3, st))

    def __call__(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._launch()
        return self.loss, self.grad


class BatchHardHostStep:
    """Host-array batch-hard step without copies: the embeddings are written straight into the library's page-locked
    staging block and the results are read out of it.

    `emb`, `labels` (and `dloss`) are numpy views of that block; `__call__` is ONE foreign call = one kernel launch + one
    wait at B <= 128 (the kernel reads and writes the block across PCIe), ~30 us host to host for 72 x 128 floats.  The
    views belong to the calling thread and stay valid until it asks the library for a larger block.

        step = BatchHardHostStep(72, 128, variant=LOSS_BH_COSINE, alpha=0.35)
        step.emb[:] = embeddings; step.labels[:] = int_labels
        loss, grad = step()            # numpy views: step.loss [B], step.grad [B, D], step.pos_idx, step.neg_idx, step.stats
    """

    def __init__(self, B: int, D: int, variant: int = _ffi.LOSS_BH_COSINE, alpha: float = 0.35, use_dloss: bool = False):
        import ctypes as C

        _ffi.init()
        self._lib = _ffi.load_library()
        self.B, self.D, self.variant, self.alpha = int(B), int(D), int(variant), float(alpha)
        bufs = (C.c_void_p * 8)()
        _ffi.check(self._lib.dif_batch_hard_host_buffers(self.B, self.D, bufs))

        def view(i, n, dtype):
            return np.ctypeslib.as_array(C.cast(bufs[i], C.POINTER(C.c_float if dtype == np.float32 else C.c_int32)), shape=(n,))

        self.emb = view(0, B * D, np.float32).reshape(B, D)
        self.labels = view(1, B, np.int32)
        self.dloss = view(2, B, np.float32) if use_dloss else None
        self.loss = view(3, B, np.float32)
        self.pos_idx, self.neg_idx = view(4, B, np.int32), view(5, B, np.int32)
        self.stats = view(6, 4, np.float32)
        self.grad = view(7, B * D, np.float32).reshape(B, D)
        self._args = (bufs[0], bufs[1], self.B, self.D, self.variant, self.alpha, bufs[3], bufs[4], bufs[5], bufs[6],
                      bufs[2] if use_dloss else None, bufs[7], _ffi.PREC_TF32X3)

    def __call__(self):
        rc = self._lib.dif_batch_hard_host(*self._args)
        if rc:
            _ffi.check(rc)
        return self.loss, self.grad


def _torch_function():
    import torch

    class _BatchHardFn(torch.autograd.Function):
        """loss [B] (differentiable) plus the by-products of the mining pass: statistics and mined columns."""

        @staticmethod
        def forward(ctx, emb, labels, variant, alpha):
            loss, _, info = batch_hard(labels, emb, variant, alpha, want_grad=False)
            ctx.save_for_backward(emb)
            ctx.labels, ctx.variant, ctx.alpha = labels, variant, alpha
            ctx.mark_non_differentiable(info["stats"], info["pos_idx"], info["neg_idx"])
            return loss, info["stats"], info["pos_idx"], info["neg_idx"]

        @staticmethod
        def backward(ctx, dloss, *_unused):
            # the gradient kernels need the cotangent, which only exists now: one fused call mines again (the B x B
            # matrix is never stored) and scatters dloss into the mined rows
            (emb,) = ctx.saved_tensors
            _, grad, _ = batch_hard(ctx.labels, emb, ctx.variant, ctx.alpha, dloss=dloss, want_grad=True)
            return grad, None, None, None

    return _BatchHardFn


class TripletLossWapper(_LossBase):
    """common/losses.py:5-30 (the reference's spelling is kept)."""

    _variant = None

    def __init__(self, alpha=0.35, soft=False, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha
        # soft=True: log(1 + exp(.)) margin of arXiv 1703.07737 eq. 4 instead of the reference's hard margin
        # (an extension - the reference only has max(. + alpha, 0), common/losses.py:51,85)
        self.soft = bool(soft)
        self.last_info = None

    def __calculate_triplet_loss__(self, y_true, y_pred, alpha):
        return None  # common/losses.py:17-18

    def call(self, labels, embeddings):
        return self.__calculate_triplet_loss__(labels, embeddings, self.alpha)

    def get_config(self):
        config = super().get_config()
        config.update({"alpha": self.alpha})
        if self.soft:
            config["soft"] = True
        return config

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    # ---- shared implementation of the subclasses
    def _dispatch(self, labels, embeddings, alpha=None):
        """Per-sample loss with margin `alpha` (None: `_current_alpha()`, which AutoAlpha maps to its state variable).
        torch-CUDA embeddings stay on the device and are differentiable; TensorFlow tensors go through
        tf.custom_gradient; anything else is treated as a host array."""
        if _tf is not None and isinstance(embeddings, (_tf.Tensor, _tf.Variable)):
            return self._tf_call(labels, embeddings, alpha)
        alpha = self._current_alpha() if alpha is None else alpha
        if _ffi.is_device_tensor(embeddings):
            loss, stats, pos, neg = _torch_function().apply(embeddings, labels, self._code(), float(alpha))
            info = {"stats": stats, "pos_idx": pos, "neg_idx": neg}
        else:
            loss, _, info = batch_hard(labels, embeddings, self._code(), alpha, want_grad=False)
        self._after_step(info)
        self.last_info = info
        return loss

    def _tf_call(self, labels, embeddings, alpha=None):
        """TensorFlow bridge (Keras `fit` in graph mode: networks/triplet.py:182,209,211, training/triplet.py:51-57):
        the kernels run inside tf.numpy_function, the gradient is attached with tf.custom_gradient.  The margin is
        read when the step RUNS (AutoAlpha's state changes between steps), and the backward pass reuses the margin
        its forward pass saw."""
        variant = self._code()
        used = {}

        @_tf.custom_gradient
        def op(emb):
            def fwd(e, l):
                used["alpha"] = float(self._current_alpha() if alpha is None else alpha)
                loss, _, info = batch_hard(l, e, variant, used["alpha"], want_grad=False)
                self._after_step(info)
                self.last_info = info
                return loss

            loss = _tf.numpy_function(fwd, [emb, labels], _tf.float32)

            def grad_fn(dloss):
                def bwd(e, l, dl):
                    _, g, _ = batch_hard(l, e, variant, used["alpha"], dloss=dl, want_grad=True)
                    return g

                return _tf.numpy_function(bwd, [emb, labels, dloss], _tf.float32)

            return loss, grad_fn

        return op(embeddings)

    def loss_and_grad(self, labels, embeddings, dloss=None):
        """(loss [B], gradient of sum_i dloss_i * loss_i (default: mean) wrt embeddings, info) in one call."""
        loss, grad, info = batch_hard(labels, embeddings, self._code(), self._current_alpha(), dloss=dloss)
        self._after_step(info)
        self.last_info = info
        return loss, grad, info

    def _code(self):
        return self._variant | (_ffi.LOSS_SOFT_MARGIN if self.soft else 0)

    def _current_alpha(self):
        return self.alpha

    def _after_step(self, info):
        pass


class BatchHardTripletLoss(TripletLossWapper):
    """common/losses.py:33-51: cosine batch-hard, hard margin."""

    _variant = _ffi.LOSS_BH_COSINE

    def __calculate_triplet_loss__(self, labels, embeddings, alpha):
        return self._dispatch(labels, embeddings, alpha)


class BatchHardTripletLossEuclidean(TripletLossWapper):
    """common/losses.py:54-85: squared-L2 batch-hard; the tf.print statistics are kept in `last_info['stats']`
    = (mean(dists), mean(hardest_pos), mean(hardest_neg), max(dists))."""

    _variant = _ffi.LOSS_BH_EUCLIDEAN

    def __calculate_triplet_loss__(self, labels, embeddings, alpha):
        return self._dispatch(labels, embeddings, alpha)


class BatchHardTripletLossEuclideanAutoAlpha(TripletLossWapper):
    """common/losses.py:88-128: the margin is the state `auto_alpha` (init 1, not trainable, not in get_config);
    each call uses the PREVIOUS value (:112) and then sets auto_alpha = mean(dists) * alpha (:113)."""

    _variant = _ffi.LOSS_BH_EUCLIDEAN

    def __init__(self, alpha=0.1, init_auto_alpha=1, **kwargs):
        super().__init__(alpha=alpha, **kwargs)
        self._mean_dists = None
        self.auto_alpha = float(init_auto_alpha)

    @property
    def auto_alpha(self) -> float:
        """The state variable of losses.py:93.  After a device-side step it is held as mean(dists) on the device and
        only read back (one 4-byte copy) when the next step needs the number - by then the kernel has long finished."""
        if self._mean_dists is not None:
            self._auto_alpha = float(self._mean_dists) * self.alpha     # :113
            self._mean_dists = None
        return self._auto_alpha

    @auto_alpha.setter
    def auto_alpha(self, value):
        self._auto_alpha, self._mean_dists = float(value), None

    def _current_alpha(self):
        return self.auto_alpha

    def _after_step(self, info):
        self._mean_dists = info["stats"][0]     # numpy scalar or 0-d device tensor: no synchronisation here

    def __calculate_triplet_loss__(self, labels, embeddings, alpha):
        return self._dispatch(labels, embeddings)     # `alpha` (the 0.1 factor) only scales the next margin


def batch_all(labels, embeddings, alpha: float, dloss=None, want_grad: bool = True):
    """Framework-neutral core of BatchAllTripletLoss (common/losses.py:131-148).  Returns (loss, grad or None)."""
    import torch

    lib = _ffi.load_library()
    on_dev = _ffi.is_device_tensor(embeddings)
    emb = embeddings.detach().contiguous().float() if on_dev else torch.from_numpy(
        _ffi.host_array(embeddings, np.float32)).cuda()
    dev = emb.device
    _ffi.init(dev.index or 0)
    B, D = emb.shape
    st = _ffi.current_stream_ptr(dev)
    if _ffi.is_device_tensor(labels) and labels.dim() == 2:
        oh = labels.detach().contiguous().float()
        lab = torch.empty(B, dtype=torch.int32, device=dev)
        _ffi.check(lib.dif_labels_from_onehot(_ffi.ptr(oh), B, oh.shape[1], _ffi.ptr(lab), st))
    elif _ffi.is_device_tensor(labels):
        lab = labels.detach().to(torch.int32).contiguous()
    else:
        lab = torch.from_numpy(_int_labels(labels)).to(dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    grad = torch.empty_like(emb) if want_grad else None
    dl = None
    if dloss is not None:
        dl = dloss.detach().contiguous().float() if _ffi.is_device_tensor(dloss) else torch.from_numpy(
            _ffi.host_array(dloss, np.float32, (B,))).to(dev)
    _ffi.check(lib.dif_batch_all(_ffi.ptr(emb), _ffi.ptr(lab), B, D, float(alpha), _ffi.ptr(loss), _ffi.ptr(dl),
                                 _ffi.ptr(grad), st))
    if on_dev:
        return loss, grad
    return loss.cpu().numpy(), (None if grad is None else grad.cpu().numpy())


class BatchAllTripletLoss(TripletLossWapper):
    """common/losses.py:131-148: cosine batch-all (mean positive distance + mean of the still-violating negatives)."""

    def __calculate_triplet_loss__(self, labels, embeddings, alpha):
        if _ffi.is_device_tensor(embeddings) and embeddings.requires_grad:
            import torch

            class _Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, emb):
                    loss, _ = batch_all(labels, emb, alpha, want_grad=False)
                    ctx.save_for_backward(emb)
                    return loss

                @staticmethod
                def backward(ctx, dloss):
                    (emb,) = ctx.saved_tensors
                    return batch_all(labels, emb, alpha, dloss=dloss, want_grad=True)[1]

            return _Fn.apply(embeddings)
        return batch_all(labels, embeddings, alpha, want_grad=False)[0]

    def loss_and_grad(self, labels, embeddings, dloss=None):
        loss, grad = batch_all(labels, embeddings, self.alpha, dloss=dloss, want_grad=True)
        return loss, grad, {}
