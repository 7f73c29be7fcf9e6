"""Drop-ins for the two tensorflow_addons losses the reference trains its triplet models with:
`tfa.losses.TripletHardLoss()` (deep_insight_face/networks/triplet.py:196,211) and
`tfa.losses.TripletSemiHardLoss()` (:209), both with their defaults and sparse integer labels
(`class_mode='sparse'`, training/triplet.py:72).

Same constructor arguments as tfa (margin=1.0, soft=False, distance_metric="L2" | "squared-L2") and the same
result: a SCALAR (tfa reduces inside the loss).  The arithmetic runs in libdif_b200.so (csrc/tfa_triplet.cu);
there is no CPU path.  distance_metric="angular" (tfa: max(1 - x^ x^T, 0) on l2-normalised rows) is served by the same
kernels: on unit rows the squared distance is 2 - 2 x^ x^T = 2 * angular, so the angular loss with margin m is HALF the
squared-L2 loss of the normalised rows with margin 2m, and its gradient goes back through dif_l2_normalize_bwd (three
launches more; with soft=True the unit rows are scaled by sqrt(1/2) instead, whose squared distances are the angular
ones).  Callables are not implemented (the reference passes neither).

    loss(y_true, y_pred)                numpy in -> python float; torch-CUDA in -> differentiable 0-d tensor
    loss.loss_and_grad(y_true, y_pred)  (loss, d loss / d y_pred [B, D], info) in one fused call
"""
from __future__ import annotations

import numpy as np

from .. import _ffi
from .losses import _LossBase

TFA_HARD, TFA_SEMIHARD, TFA_SOFT, TFA_SQUARED = 0, 1, 4, 8
TFA_ANGULAR = 16   # host-side flag: never reaches the C ABI (angular_via_squared turns it into TFA_SQUARED on unit rows)


def _metric_flag(distance_metric) -> int:
    if distance_metric == "L2":
        return 0
    if distance_metric == "squared-L2":
        return TFA_SQUARED
    if distance_metric == "angular":
        return TFA_ANGULAR
    raise NotImplementedError(f"distance_metric {distance_metric!r}: only 'L2', 'squared-L2' and 'angular' run on the GPU path")


def angular_via_squared(kind: int, margin: float, dloss: float, normalize, squared_loss, normalize_bwd):
    """tfa's angular metric out of the squared-L2 step on unit rows.  With x^ = l2_normalize(x):
    angular_ij = max(1 - x^_i x^_j, 0) = |x^_i - x^_j|^2 / 2, and every comparison the mining makes is unchanged by
    the factor 2.  Hinge losses: hinge(a_p - a_n + m) = hinge(s_p - s_n + 2m) / 2, so
        loss(x; angular, m) = 0.5 * loss(x^; squared-L2, 2m),   d loss / d x = J_normalize^T (0.5 * d loss_sq / d x^)
    with both factors exact in fp32.  soft=True (log1p(exp(a_p - a_n)) does not commute with a factor): the unit rows
    are scaled by sqrt(1/2) instead, whose squared distances ARE the angular ones (one rounding per element),
        loss(x; angular, soft) = loss(x^ * sqrt(1/2); squared-L2, soft),   d / d x^ = sqrt(1/2) * d / d (x^ sqrt(1/2)).
    The three steps are passed in (the CUDA entry points in tfa_triplet; CPU stand-ins in tests/test_oracle_cpu.py):
        normalize()                         -> (unit rows, saved state)
        squared_loss(rows, code, m, dloss)  -> (loss, d loss / d rows or None, info)
        normalize_bwd(g, unit, state)       -> d / d x
    Returns (loss, grad or None, info).  Known divergence: a row with |x|^2 < 1e-12 stays (near) zero under
    tf.math.l2_normalize, so it is not a unit row - tfa gives it the angular distance 1 to every sample, this path 1/2."""
    code = (kind & ~TFA_ANGULAR) | TFA_SQUARED
    unit, state = normalize()
    if kind & TFA_SOFT:
        r = float(np.sqrt(0.5))
        loss, g, info = squared_loss(unit * r, code, float(margin), float(dloss))
        return loss, (None if g is None else normalize_bwd(g * r, unit, state)), info
    loss2, g2, info = squared_loss(unit, code, 2.0 * float(margin), 0.5 * float(dloss))
    return 0.5 * loss2, (None if g2 is None else normalize_bwd(g2, unit, state)), info


class TfaTripletStep:
    """Preallocated, optionally CUDA-graphed tfa hard / semi-hard step for a fixed (B, D): the training-loop form.

    At the reference's batch size (18 identities x 4 images) the step is two kernel launches of a few microseconds;
    building the output tensors and a 12-argument foreign call per step costs more than that, so the buffers are made
    once and a step is one foreign call - or, with `graph=True`, one cudaGraphLaunch.

        step = TfaTripletStep(B, D, kind=TFA_SEMIHARD, margin=1.0, device="cuda:0", graph=True)
        step.emb.copy_(embeddings); step.labels.copy_(sparse_int_labels)
        loss, grad = step()            # step.loss [1], step.grad [B, D] (d loss / d emb), step.pos_idx / neg_idx (hard)
    """

    def __init__(self, B: int, D: int, kind: int = 0, margin: float = 1.0, device="cuda:0", graph: bool = False,
                 want_grad: bool = True, dloss: float = 1.0):
        import torch

        dev = torch.device(device)
        _ffi.init(dev.index or 0)
        self._lib = _ffi.load_library()
        if int(kind) & TFA_ANGULAR:
            raise NotImplementedError("TfaTripletStep runs the 'L2' / 'squared-L2' metrics; for 'angular' normalise the "
                                      "batch (networks.head.l2_normalize) and step on squared-L2 with twice the margin")
        self.B, self.D, self.kind, self.margin, self.dloss = int(B), int(D), int(kind), float(margin), float(dloss)
        self.emb = torch.zeros((B, D), dtype=torch.float32, device=dev)
        self.labels = torch.zeros(B, dtype=torch.int32, device=dev)
        self.loss = torch.empty(1, dtype=torch.float32, device=dev)
        hard = (self.kind & 3) == TFA_HARD
        self.pos_idx = torch.empty(B, dtype=torch.int32, device=dev) if hard else None
        self.neg_idx = torch.empty(B, dtype=torch.int32, device=dev) if hard else None
        self.grad = torch.empty((B, D), dtype=torch.float32, device=dev) if want_grad else None
        self._dev = dev
        self._graph = None
        self.labels.copy_(torch.arange(B, dtype=torch.int32, device=dev) // 2)   # a valid batch for the warm-up call
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
        _ffi.check(self._lib.dif_tfa_triplet(
            self.emb.data_ptr(), self.labels.data_ptr(), self.B, self.D, self.kind, self.margin, self.loss.data_ptr(),
            None if self.pos_idx is None else self.pos_idx.data_ptr(), None if self.neg_idx is None else self.neg_idx.data_ptr(),
            self.dloss, None if self.grad is None else self.grad.data_ptr(), st))

    def __call__(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._launch()
        return self.loss, self.grad


def tfa_triplet(labels, embeddings, kind: int, margin: float = 1.0, dloss: float = 1.0, want_grad: bool = True,
                want_indices: bool = True):
    """Framework-neutral core.  Returns (loss, grad or None, info); device tensors in -> device tensors out."""
    import torch

    lib = _ffi.load_library()
    on_dev = _ffi.is_device_tensor(embeddings)
    emb = embeddings.detach().contiguous().float() if on_dev else torch.from_numpy(
        _ffi.host_array(embeddings, np.float32)).cuda()
    if emb.dim() != 2:
        raise ValueError("embeddings must be [B, D]")
    if kind & TFA_ANGULAR:
        from ..networks.head import _launch_bwd, _launch_fwd

        dev_labels = labels if _ffi.is_device_tensor(labels) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(_ffi.host_array(labels, None)).reshape(-1), dtype=np.int32)).to(emb.device)
        loss, grad, info = angular_via_squared(
            kind, margin, dloss, lambda: _launch_fwd(emb),
            lambda unit, code, m, dl: tfa_triplet(dev_labels, unit, code, m, dloss=dl, want_grad=want_grad,
                                                  want_indices=want_indices),
            lambda g, unit, inv: _launch_bwd(g, unit, inv))
        if on_dev:
            return loss, grad, info
        info = {k: v.cpu().numpy() for k, v in info.items()}
        return float(loss.cpu()), (None if grad is None else grad.cpu().numpy()), info
    dev = emb.device
    _ffi.init(dev.index or 0)
    B, D = emb.shape
    st = _ffi.current_stream_ptr(dev)
    if _ffi.is_device_tensor(labels):
        lab = labels.detach().reshape(-1).to(torch.int32).contiguous()
    else:
        lab_np = np.asarray(_ffi.host_array(labels, None))
        if lab_np.ndim == 2 and lab_np.shape[1] == 1:   # tfa reshapes [B] or [B, 1] to a column
            lab_np = lab_np.reshape(-1)
        if lab_np.ndim != 1:
            raise ValueError("labels must be sparse integer class ids [B] (tfa convention), not one-hot")
        lab = torch.from_numpy(np.ascontiguousarray(lab_np, dtype=np.int32)).to(dev)
    if lab.shape != (B,):
        raise ValueError(f"labels must be sparse int [B] (or [B, 1]); got {tuple(lab.shape)} for B={B}")
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    hard = (kind & 3) == TFA_HARD
    pos = torch.empty(B, dtype=torch.int32, device=dev) if (hard and want_indices) else None
    neg = torch.empty(B, dtype=torch.int32, device=dev) if (hard and want_indices) else None
    grad = torch.empty_like(emb) if want_grad else None
    _ffi.check(lib.dif_tfa_triplet(_ffi.ptr(emb), _ffi.ptr(lab), B, D, int(kind), float(margin), _ffi.ptr(loss),
                                   _ffi.ptr(pos), _ffi.ptr(neg), float(dloss), _ffi.ptr(grad), st))
    info = {} if pos is None else {"pos_idx": pos, "neg_idx": neg}
    if on_dev:
        return loss[0], grad, info
    info = {k: v.cpu().numpy() for k, v in info.items()}
    return float(loss.cpu()[0]), (None if grad is None else grad.cpu().numpy()), info


class _TfaTripletBase(_LossBase):
    _kind = TFA_HARD

    def __init__(self, margin=1.0, distance_metric="L2", name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.margin = margin
        self.distance_metric = distance_metric
        self._flag = _metric_flag(distance_metric)
        self.last_info = None

    def _code(self) -> int:
        return self._kind | self._flag

    def call(self, y_true, y_pred):
        if _ffi.is_device_tensor(y_pred) and y_pred.requires_grad:
            import torch

            code, margin = self._code(), self.margin

            class _Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, emb):
                    loss, _, _ = tfa_triplet(y_true, emb, code, margin, want_grad=False, want_indices=False)
                    ctx.save_for_backward(emb)
                    return loss

                @staticmethod
                def backward(ctx, dloss):
                    (emb,) = ctx.saved_tensors
                    return tfa_triplet(y_true, emb, code, margin, dloss=float(dloss), want_indices=False)[1]

            return _Fn.apply(y_pred)
        loss, _, info = tfa_triplet(y_true, y_pred, self._code(), self.margin, want_grad=False)
        self.last_info = info
        return loss

    def __call__(self, y_true, y_pred, sample_weight=None):
        return self.call(y_true, y_pred)   # already a scalar, as in tfa

    def loss_and_grad(self, y_true, y_pred, dloss: float = 1.0):
        loss, grad, info = tfa_triplet(y_true, y_pred, self._code(), self.margin, dloss=dloss)
        self.last_info = info
        return loss, grad, info

    def get_config(self):
        config = super().get_config()
        config.update({"margin": self.margin, "distance_metric": self.distance_metric})
        return config

    @classmethod
    def from_config(cls, config):
        return cls(**config)


class TripletHardLoss(_TfaTripletBase):
    """tfa.losses.TripletHardLoss(margin=1.0, soft=False, distance_metric="L2"): mean over anchors of
    max(hardest positive - hardest negative + margin, 0), or log1p(exp(.)) when soft."""

    _kind = TFA_HARD

    def __init__(self, margin=1.0, soft=False, distance_metric="L2", name=None, **kwargs):
        super().__init__(margin=margin, distance_metric=distance_metric, name=name, **kwargs)
        self.soft = bool(soft)

    def _code(self) -> int:
        return super()._code() | (TFA_SOFT if self.soft else 0)

    def get_config(self):
        config = super().get_config()
        config["soft"] = self.soft
        return config


class TripletSemiHardLoss(_TfaTripletBase):
    """tfa.losses.TripletSemiHardLoss(margin=1.0, distance_metric="L2"): for every positive pair the closest
    negative farther than the positive (else the farthest negative); sum of the hinges / number of pairs."""

    _kind = TFA_SEMIHARD
