"""The Python glue of the tfa drop-ins, end to end on the CPU against a FAKE library: every C-ABI entry point the
path calls (dif_tfa_triplet, dif_l2_normalize, dif_l2_normalize_bwd) is replaced by a stand-in that reads and writes
the caller's buffers through the raw pointers it is handed and computes with the oracle.  What is checked is the
wiring - argument order, flags, buffers, the numpy / device-tensor / autograd routes - not the kernels (those are the
-m gpu tests).  The oracle is the checker's arithmetic here, never the product's: the product code under test is the
glue, and it cannot tell the fake from libdif_b200.so."""
import ctypes as C

import numpy as np
import pytest


def _view(ptr, shape, ctype=C.c_float):
    return np.ctypeslib.as_array(C.cast(int(ptr), C.POINTER(ctype)), shape=tuple(shape))


class FakeLib:
    """Stand-ins with the signatures of include/dif_b200.h; `calls` records (entry point, kind, margin, dloss)."""

    def __init__(self):
        self.calls = []

    def dif_l2_normalize(self, x, n, D, y, inv, stream):
        xv = _view(x, (n, D)).astype(np.float64)
        r = 1.0 / np.sqrt(np.maximum((xv * xv).sum(1), 1e-12))
        _view(y, (n, D))[:] = xv * r[:, None]
        _view(inv, (n,))[:] = r
        self.calls.append(("l2_normalize",))
        return 0

    def dif_l2_normalize_bwd(self, g, y, inv, n, D, dx, stream):
        gv, yv, r = _view(g, (n, D)).astype(np.float64), _view(y, (n, D)).astype(np.float64), _view(inv, (n,)).astype(np.float64)
        _view(dx, (n, D))[:] = r[:, None] * (gv - yv * (yv * gv).sum(1, keepdims=True))
        self.calls.append(("l2_normalize_bwd",))
        return 0

    def dif_tfa_triplet(self, emb, lab, B, D, kind, margin, loss, pos, neg, dloss, grad, stream):
        from oracle import tfa_oracle as t

        assert kind & ~(1 | 4 | 8) == 0, f"flag bits the C ABI does not know: {kind}"
        x, labels = _view(emb, (B, D)).copy(), _view(lab, (B,), C.c_int32).copy()
        name, squared, soft = ("semihard" if kind & 1 else "hard"), bool(kind & 8), bool(kind & 4)
        import torch

        with torch.enable_grad():     # the oracle differentiates with autograd; the glue may call from a no-grad forward
            l, g = t.torch_shadow(name, labels, x, margin=margin, soft=soft, squared=squared)
        _view(loss, (1,))[0] = l
        if grad:
            _view(grad, (B, D))[:] = dloss * g
        if pos and neg:
            h = t.triplet_hard(labels, x, margin=margin, soft=soft, squared=squared)
            _view(pos, (B,), C.c_int32)[:] = h["pos_idx"]
            _view(neg, (B,), C.c_int32)[:] = h["neg_idx"]
        self.calls.append(("tfa_triplet", kind, margin, dloss))
        return 0


@pytest.fixture()
def fake(monkeypatch):
    import torch

    from deep_insight_face_b200 import _ffi

    lib = FakeLib()
    monkeypatch.setattr(_ffi, "load_library", lambda: lib)
    monkeypatch.setattr(_ffi, "init", lambda device=None: None)
    monkeypatch.setattr(_ffi, "current_stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    # every torch tensor plays a device tensor (numpy inputs stay host arrays, as on the GPU box)
    monkeypatch.setattr(_ffi, "is_device_tensor", lambda a: isinstance(a, torch.Tensor))
    return lib


def _pk(P, K, D, seed):
    rng = np.random.default_rng(seed)
    lab = np.repeat(np.arange(P), K).astype(np.int32)
    x = (rng.normal(size=(P, D))[lab] + 0.7 * rng.normal(size=(P * K, D))).astype(np.float32) * 3.0
    perm = rng.permutation(P * K)
    return lab[perm], x[perm]


@pytest.mark.parametrize("kind", ["hard", "semihard"])
def test_angular_route_numpy_and_autograd(fake, kind):
    import torch

    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(18, 4, 64, seed=5)
    margin = 0.7
    fn = (TripletHardLoss if kind == "hard" else TripletSemiHardLoss)(margin=margin, distance_metric="angular")
    want, want_g = orc.torch_shadow_angular_fp64(kind, lab, x, margin)
    scale = np.abs(want_g).max()
    assert want > 0.05 and scale > 0

    loss, grad, info = fn.loss_and_grad(lab, x)                  # numpy in -> numpy out
    base = 1 if kind == "semihard" else 0
    assert fake.calls == [("l2_normalize",), ("tfa_triplet", base | 8, 2.0 * margin, 0.5), ("l2_normalize_bwd",)]
    assert isinstance(loss, float) and abs(loss - want) <= 1e-5 * max(1.0, abs(want))
    assert isinstance(grad, np.ndarray) and np.abs(grad - want_g).max() <= 1e-4 * scale
    if kind == "hard":
        unit = x / np.linalg.norm(x, axis=1, keepdims=True)
        h = orc.triplet_hard(lab, unit.astype(np.float32), margin=2 * margin, squared=True)
        assert np.array_equal(info["pos_idx"], h["pos_idx"]) and np.array_equal(info["neg_idx"], h["neg_idx"])

    fake.calls.clear()
    assert abs(fn(lab, x) - want) <= 1e-5 * max(1.0, abs(want))    # loss only: no backward launches
    assert [c[0] for c in fake.calls] == ["l2_normalize", "tfa_triplet"]

    # device-tensor route (CPU tensors passed off as device tensors): differentiable 0-d tensor
    e = torch.from_numpy(x.copy()).requires_grad_(True)
    out = fn(torch.from_numpy(lab), e)
    assert out.dim() == 0 and abs(float(out.detach()) - want) <= 1e-5 * max(1.0, abs(want))
    (2.0 * out).backward()
    assert np.abs(e.grad.numpy() - 2.0 * want_g).max() <= 2e-4 * scale


@pytest.mark.parametrize("metric,flag", [("L2", 0), ("squared-L2", 8)])
def test_plain_routes_pass_the_flags_through(fake, metric, flag):
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(9, 3, 32, seed=2)
    x = x * 0.2
    for cls, kind, extra in ((TripletHardLoss, "hard", 0), (TripletSemiHardLoss, "semihard", 1)):
        fake.calls.clear()
        loss, grad, _ = cls(margin=0.9, distance_metric=metric).loss_and_grad(lab, x, dloss=3.0)
        assert fake.calls == [("tfa_triplet", extra | flag, 0.9, 3.0)]
        l, g = orc.torch_shadow(kind, lab, x, margin=0.9, squared=bool(flag))
        assert abs(loss - l) <= 1e-6 * max(1.0, abs(l)) and np.abs(grad - 3.0 * g).max() <= 1e-5 * max(np.abs(g).max(), 1e-6) * 3
    fake.calls.clear()
    TripletHardLoss(soft=True, distance_metric=metric).loss_and_grad(lab, x)
    assert fake.calls[0][1] == 4 | flag


def test_soft_angular_route_scales_the_unit_rows(fake):
    """soft=True with 'angular': unit rows times sqrt(1/2) through the squared-L2 soft loss, margin and cotangent as
    they are, gradient times sqrt(1/2) back through the normalisation."""
    import torch

    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(12, 4, 48, seed=9)
    fn = TripletHardLoss(soft=True, distance_metric="angular")
    want, want_g = orc.torch_shadow_angular_fp64("hard", lab, x, 1.0, soft=True)
    scale = np.abs(want_g).max()
    loss, grad, _ = fn.loss_and_grad(lab, x)
    assert fake.calls == [("l2_normalize",), ("tfa_triplet", 4 | 8, 1.0, 1.0), ("l2_normalize_bwd",)]
    assert abs(loss - want) <= 1e-5 * max(1.0, abs(want)) and np.abs(grad - want_g).max() <= 1e-4 * scale
    e = torch.from_numpy(x.copy()).requires_grad_(True)
    out = fn(torch.from_numpy(lab), e)
    out.backward()
    assert abs(float(out.detach()) - want) <= 1e-5 * max(1.0, abs(want))
    assert np.abs(e.grad.numpy() - want_g).max() <= 1e-4 * scale

