"""The host logic of the Keras-facing loss classes (common/losses.py mirror of the reference's common/losses.py:5-128)
on the CPU, against a FAKE library: dif_batch_hard / dif_batch_hard_host / dif_labels_from_onehot are stand-ins that
read and write the caller's buffers through the raw pointers they are handed and compute with the oracle.  Checked here:
which margin reaches the library at each step (AutoAlpha's previous-value semantics, losses.py:112-113), the label
layouts, shapes and reductions, get_config / from_config, the torch autograd route and the TensorFlow bridge under a
stand-in module.  The arithmetic of the kernels is the -m gpu tests' business; the product code under test is the
glue, which cannot tell the fake from libdif_b200.so."""
import ctypes as C

import numpy as np
import pytest

from test_losses_gpu import _FakeTensor, _FakeTF, close, pk_batch


def _view(ptr, shape, ctype=C.c_float):
    return np.ctypeslib.as_array(C.cast(int(ptr), C.POINTER(ctype)), shape=tuple(shape))


class FakeLib:
    def __init__(self):
        self.calls = []     # (entry point, variant, alpha, has dloss, wants gradient)

    def _step(self, name, emb, lab, B, D, variant, alpha, loss, pos, neg, stats, dloss, demb):
        from oracle import losses_oracle as lo

        x, labels = _view(emb, (B, D)).copy(), _view(lab, (B,), C.c_int32).copy()
        dl = _view(dloss, (B,)).copy() if dloss else None
        fn = lo.batch_hard_euclidean if variant & 1 else lo.batch_hard_cosine
        r = fn(labels, x, alpha, dloss=dl, soft=bool(variant & 4))
        _view(loss, (B,))[:] = r["loss"]
        _view(pos, (B,), C.c_int32)[:] = r["pos_idx"]
        _view(neg, (B,), C.c_int32)[:] = r["neg_idx"]
        _view(stats, (4,))[:] = r["stats"]
        if demb:
            _view(demb, (B, D))[:] = r["grad"]
        self.calls.append((name, variant, float(alpha), dloss is not None and bool(dloss), bool(demb)))
        return 0

    def dif_batch_hard_host(self, emb, lab, B, D, variant, alpha, loss, pos, neg, stats, dloss, demb, precision):
        return self._step("host", emb, lab, B, D, variant, alpha, loss, pos, neg, stats, dloss, demb)

    def dif_batch_hard(self, emb, lab, B, D, variant, alpha, loss, pos, neg, stats, dloss, demb, precision, stream):
        return self._step("device", emb, lab, B, D, variant, alpha, loss, pos, neg, stats, dloss, demb)

    def dif_labels_from_onehot(self, onehot, B, Cn, lab, stream):
        _view(lab, (B,), C.c_int32)[:] = _view(onehot, (B, Cn)).argmax(1)
        self.calls.append(("onehot", Cn))
        return 0


@pytest.fixture()
def fake(monkeypatch):
    import torch

    from deep_insight_face_b200 import _ffi

    lib = FakeLib()
    monkeypatch.setattr(_ffi, "load_library", lambda: lib)
    monkeypatch.setattr(_ffi, "init", lambda device=None: None)
    monkeypatch.setattr(_ffi, "current_stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(_ffi, "is_device_tensor", lambda a: isinstance(a, torch.Tensor))   # torch tensors play device tensors
    return lib


def test_keras_protocol_and_auto_alpha_host_logic(fake):
    from deep_insight_face_b200.common.losses import (BatchHardTripletLoss, BatchHardTripletLossEuclidean,
                                                      BatchHardTripletLossEuclideanAutoAlpha, TripletLossWapper)
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    onehot = np.eye(18, dtype=np.float32)[lab]
    loss = BatchHardTripletLoss(alpha=0.2)
    assert loss.get_config()["alpha"] == 0.2 and "soft" not in loss.get_config()
    again = BatchHardTripletLoss.from_config(loss.get_config())
    per_sample = again.call(onehot, emb)                                   # one-hot labels, as Keras feeds them (:35)
    assert per_sample.shape == (72,) and fake.calls[-1] == ("host", 0, 0.2, False, False)
    close(per_sample, lo.batch_hard_cosine(lab, emb, 0.2)["loss"])
    close(again(onehot, emb), lo.batch_hard_cosine(lab, emb, 0.2)["loss"].mean())      # Keras AUTO reduction
    close(again.call(lab.astype(np.int64), emb), per_sample)                # sparse labels are accepted too
    assert TripletLossWapper().call(onehot, emb) is None                    # losses.py:17-18
    assert BatchHardTripletLossEuclidean(alpha=3.0, soft=True).get_config()["soft"] is True
    BatchHardTripletLossEuclidean(alpha=3.0, soft=True).call(onehot, emb)
    assert fake.calls[-1] == ("host", 1 | 4, 3.0, False, False)
    got, grad, info = BatchHardTripletLossEuclidean(alpha=3.0).loss_and_grad(onehot, emb, dloss=np.ones(72, np.float32))
    assert fake.calls[-1] == ("host", 1, 3.0, True, True) and grad.shape == emb.shape and info["stats"].shape == (4,)

    auto = BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    assert "auto_alpha" not in auto.get_config()
    auto.call(onehot, emb)
    assert fake.calls[-1][:3] == ("host", 1, 1.0)                           # the PREVIOUS value: the initial 1 (:112)
    mean_dists = float(lo.batch_hard_euclidean(lab, emb, 1.0)["stats"][0])
    assert abs(auto.auto_alpha - 0.1 * mean_dists) <= 1e-6 * mean_dists       # :113
    auto.call(onehot, 2.0 * emb)
    assert abs(fake.calls[-1][2] - 0.1 * mean_dists) <= 1e-6 * mean_dists   # step 2 runs with step 1's margin
    assert abs(auto.auto_alpha - 0.4 * mean_dists) <= 1e-5 * mean_dists       # squared distances of 2x are 4x
    with pytest.raises(ValueError):
        again.call(onehot[:10], emb)


def test_torch_autograd_route_mines_with_the_forward_margin(fake):
    import torch

    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclideanAutoAlpha
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    x = torch.from_numpy(emb.copy()).requires_grad_(True)
    onehot = torch.from_numpy(np.eye(18, dtype=np.float32)[lab])
    per_sample = BatchHardTripletLoss().call(onehot, x)
    assert fake.calls[-2:] == [("onehot", 18), ("device", 0, 0.35, False, False)]
    per_sample.mean().backward()
    assert fake.calls[-1] == ("device", 0, 0.35, True, True)
    close(x.grad.numpy(), lo.batch_hard_cosine(lab, emb, 0.35)["grad"])

    auto = BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    x1 = torch.from_numpy(emb.copy()).requires_grad_(True)
    first = auto.call(onehot, x1)
    assert first.requires_grad                                               # differentiable through call()
    second = auto.call(onehot, torch.from_numpy(emb.copy()))                 # the state moves on before step 1's backward
    margin2 = fake.calls[-1][2]
    assert abs(margin2 - 0.1 * float(lo.batch_hard_euclidean(lab, emb, 1.0)["stats"][0])) <= 1e-5 * margin2 and not second.requires_grad
    first.sum().backward()
    assert fake.calls[-1] == ("device", 1, 1.0, True, True)                  # backward of step 1 still uses margin 1
    close(x1.grad.numpy(), lo.batch_hard_euclidean(lab, emb, 1.0, dloss=np.ones(72, np.float32))["grad"])


def test_tf_bridge_host_logic(fake, monkeypatch):
    """_tf_call (tf.custom_gradient + tf.numpy_function; the path Keras `fit` takes, reference
    networks/triplet.py:182,209,211) under the stand-in module of tests/test_losses_gpu.py."""
    from deep_insight_face_b200.common import losses as L
    from oracle import losses_oracle as lo

    tf = _FakeTF()
    monkeypatch.setattr(L, "_tf", tf)
    emb, lab = pk_batch(18, 4, 128, 1.5)
    onehot = np.eye(18, dtype=np.float32)[lab]
    out = L.BatchHardTripletLoss(alpha=0.2).call(_FakeTensor(onehot), _FakeTensor(emb))
    assert isinstance(out, _FakeTensor) and tf.numpy_calls == 1 and fake.calls[-1] == ("host", 0, 0.2, False, False)
    dl = np.full(72, 1.0 / 72, dtype=np.float32)
    g = tf.grad_fns[-1](_FakeTensor(dl))
    assert fake.calls[-1] == ("host", 0, 0.2, True, True)
    close(g.numpy(), lo.batch_hard_cosine(lab, emb, 0.2)["grad"])
    auto = L.BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    auto.call(_FakeTensor(onehot), _FakeTensor(emb))
    auto.call(_FakeTensor(onehot), _FakeTensor(emb))
    tf.grad_fns[-2](_FakeTensor(dl))                                         # backward of step 1 after step 2 ran
    assert fake.calls[-1][:3] == ("host", 1, 1.0)
