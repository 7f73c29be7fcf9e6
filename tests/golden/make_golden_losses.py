"""Generates tests/golden/losses_reference.npz by EXECUTING THE REFERENCE'S OWN LOSS SOURCE, in the build container:

    python tests/golden/make_golden_losses.py

TensorFlow / Keras cannot be installed here, so the reference code runs on a stand-in: a module named `tensorflow`
(and a `K` backend object) whose two dozen entry points - exactly the ones the loss code calls - are one-line torch
expressions in float64 with autograd.  What is executed is the reference's text, not a restatement of it:

 * deep_insight_face/common/losses.py is IMPORTED as a module (its only imports are `tensorflow` and
   `tensorflow.python.ops`): BatchHardTripletLoss, BatchHardTripletLossEuclidean,
   BatchHardTripletLossEuclideanAutoAlpha, BatchAllTripletLoss are instantiated and `call(labels, embeddings)`ed;
 * networks/triplet.py:triplet_loss and networks/siamese.py:{euclidean_distance, contrastive_loss, _accuracy} live in
   modules that import the whole Keras model zoo, so their source is cut out with `ast` and executed as is;
 * networks/utils.py (numpy only) is imported; api.py:{face_distance, compare_faces} are cut out and executed.

The stand-in follows TensorFlow where the two libraries differ: `maximum(x, y)` sends the gradient to x where x >= y
(torch.maximum splits it at ties), `reduce_min / reduce_max` split the cotangent evenly over tied positions (torch.amin /
amax do the same), `l2_normalize` is x * rsqrt(max(sum x^2, 1e-12)).  Gradients are those of mean(loss), the
reduction Keras applies.  /root/reference does not exist on the GPU box; only the .npz travels.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DT = torch.float64


# ------------------------------------------------------------------ the stand-in
def _t(x):
    return x.value if isinstance(x, Variable) else x


class Variable:
    """tf.Variable as the loss code uses it: a scalar that takes part in arithmetic and can be assigned."""

    def __init__(self, initial_value, dtype=None, trainable=False):
        self.value = torch.tensor(float(initial_value), dtype=DT)

    def assign(self, v):
        self.value = _t(v).detach().clone()

    def __add__(self, o):
        return self.value + _t(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.value - _t(o)

    def __rsub__(self, o):
        return _t(o) - self.value

    def __mul__(self, o):
        return self.value * _t(o)

    __rmul__ = __mul__


def _maximum(x, y):
    x = _t(x)
    y = torch.as_tensor(_t(y), dtype=x.dtype)
    return torch.where(x >= y, x, y.expand_as(x))          # tf.maximum: the gradient goes to x where x >= y


def _reduce(fn_all, fn_dim):
    def f(x, axis=None, keepdims=False):
        x = _t(x)
        return fn_all(x) if axis is None else fn_dim(x, axis, keepdims)
    return f


class _Loss:
    """tf.keras.losses.Loss: only what TripletLossWapper touches."""

    def __init__(self, reduction="auto", name=None, **kwargs):
        self.reduction, self.name = reduction, name

    def get_config(self):
        return {"reduction": self.reduction, "name": self.name}


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32 = DT                                         # everything is computed in float64
    tf.argmax = lambda x, axis=None: torch.argmax(_t(x), dim=axis)
    tf.equal = lambda a, b: _t(a) == _t(b)
    tf.expand_dims = lambda x, axis: torch.unsqueeze(_t(x), axis)
    tf.matmul = lambda a, b: _t(a) @ _t(b)
    tf.transpose = lambda x: _t(x).t()
    tf.where = lambda c, a, b: torch.where(c, _t(a), _t(b))
    tf.ones_like = lambda x: torch.ones_like(_t(x))
    tf.zeros_like = lambda x: torch.zeros_like(_t(x))
    tf.square = lambda x: _t(x) * _t(x)
    tf.reshape = lambda x, shape: torch.reshape(_t(x), shape)
    tf.maximum = _maximum
    tf.logical_and = lambda a, b: a & b
    tf.logical_not = lambda a: ~a
    tf.cast = lambda x, dtype=None: _t(x).to(dtype)
    tf.reduce_min = _reduce(torch.amin, lambda x, a, k: torch.amin(x, dim=a, keepdim=k))
    tf.reduce_max = _reduce(lambda x: torch.amax(x.reshape(-1), dim=0), lambda x, a, k: torch.amax(x, dim=a, keepdim=k))
    tf.reduce_sum = _reduce(torch.sum, lambda x, a, k: torch.sum(x, dim=a, keepdim=k))
    tf.reduce_mean = _reduce(torch.mean, lambda x, a, k: torch.mean(x, dim=a, keepdim=k))
    tf.print = lambda *a, **k: None
    tf.Variable = Variable
    tf.Tensor = torch.Tensor
    nn = types.ModuleType("tensorflow.nn")
    nn.l2_normalize = lambda x, axis, epsilon=1e-12: _t(x) * torch.rsqrt(
        torch.clamp_min(torch.sum(_t(x) * _t(x), dim=axis, keepdim=True), epsilon))
    tf.nn = nn
    keras = types.ModuleType("tensorflow.keras")
    losses = types.ModuleType("tensorflow.keras.losses")
    losses.Loss = _Loss
    keras.losses = losses
    tf.keras = keras
    python = types.ModuleType("tensorflow.python")
    ops = types.ModuleType("tensorflow.python.ops")
    ops.array_ops = types.ModuleType("array_ops")          # imported by the file, used only by the dead code of a6
    ops.math_ops = types.ModuleType("math_ops")
    python.ops = ops
    tf.python = python
    return tf, {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.keras": keras, "tensorflow.keras.losses": losses,
                "tensorflow.python": python, "tensorflow.python.ops": ops}


class _K:
    """keras.backend as networks/triplet.py and networks/siamese.py use it."""

    @staticmethod
    def sum(x, axis=None, keepdims=False):
        return torch.sum(x, dim=axis, keepdim=keepdims)

    @staticmethod
    def square(x):
        return x * x

    maximum = staticmethod(_maximum)

    @staticmethod
    def sqrt(x):
        return torch.sqrt(x)

    @staticmethod
    def epsilon():
        return 1e-7

    @staticmethod
    def mean(x):
        return torch.mean(x.to(DT))

    @staticmethod
    def cast(x, dtype):
        return x.to(dtype)

    @staticmethod
    def equal(a, b):
        return a == b


class _Pred(torch.Tensor):
    """y_pred of triplet_loss: the function asks for `y_pred.shape.as_list()`."""

    class _Shape(tuple):
        def as_list(self):
            return list(self)

    @property
    def shape(self):
        return _Pred._Shape(super().shape)


def function_source(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    return ast.get_source_segment(src, node)


# ------------------------------------------------------------------ inputs (regenerated identically by the tests)
def pk_batch(P, K, D, noise, seed, duplicates=False, zero_row=False, one_identity=False):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((P, D)).astype(np.float32)
    lab = np.repeat(np.arange(P), K)
    emb = (cent[lab] + noise * rng.standard_normal((P * K, D))).astype(np.float32)
    perm = rng.permutation(P * K)
    emb, lab = emb[perm], lab[perm]
    if duplicates:
        emb[3] = emb[1]
        emb[P * K - 1] = emb[0]
    if zero_row:
        emb[5] = 0.0
    if one_identity:
        lab[:] = 0
    return emb, lab.astype(np.int64)


CASES = [  # name, P, K, D, noise, seed, flags
    ("c1", 18, 4, 128, 1.0, 0, {}),
    ("c1_tight", 18, 4, 128, 0.3, 1, {}),
    ("wide", 33, 3, 100, 1.5, 2, {}),
    ("dups", 8, 4, 64, 0.7, 3, {"duplicates": True}),
    ("zero", 12, 3, 32, 1.0, 4, {"zero_row": True}),
    ("one_identity", 6, 4, 48, 1.0, 5, {"one_identity": True}),
    ("singletons", 24, 1, 64, 1.0, 6, {}),
    ("big", 128, 4, 128, 1.0, 7, {}),
]


def run_loss(obj, emb, lab, n_classes):
    x = torch.tensor(emb.astype(np.float64), requires_grad=True)
    onehot = torch.tensor(np.eye(n_classes)[lab])
    loss = obj.call(onehot, x)                               # the reference's code
    loss.mean().backward()
    return loss.detach().numpy(), x.grad.numpy()


def main():
    tf, mods = make_tf()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(REF, "deep_insight_face/common/losses.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)                         # the reference's module, unmodified
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    out = {}
    for name, P, K, D, noise, seed, flags in CASES:
        emb, lab = pk_batch(P, K, D, noise, seed, **flags)
        n_cls = int(lab.max()) + 1
        for key, obj in (("bh_cos", ref.BatchHardTripletLoss(alpha=0.35)),
                         ("bh_euc", ref.BatchHardTripletLossEuclidean(alpha=0.3 * D)),
                         ("ball", ref.BatchAllTripletLoss(alpha=0.35))):
            loss, grad = run_loss(obj, emb, lab, n_cls)
            out[f"{name}/{key}/loss"], out[f"{name}/{key}/grad"] = loss, grad
        auto = ref.BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
        for step in range(2):                                # step 0 uses the initial margin 1, step 1 mean(dists) * alpha
            loss, grad = run_loss(auto, emb, lab, n_cls)
            out[f"{name}/bh_auto{step}/loss"], out[f"{name}/bh_auto{step}/grad"] = loss, grad
            out[f"{name}/bh_auto{step}/auto_alpha_after"] = np.float64(auto.auto_alpha.value.item())
    # networks/triplet.py:16-46 and networks/siamese.py:22-45, source cut out and executed
    ns = {"K": _K, "tf": tf}
    exec(function_source(os.path.join(REF, "deep_insight_face/networks/triplet.py"), "triplet_loss"), ns)
    for fn in ("euclidean_distance", "contrastive_loss", "_accuracy"):
        exec(function_source(os.path.join(REF, "deep_insight_face/networks/siamese.py"), fn), ns)
    rng = np.random.default_rng(11)
    for name, B, D in (("apn_a", 64, 128), ("apn_b", 37, 48)):
        y = rng.standard_normal((B, 3 * D)).astype(np.float32)
        y[:, D:2 * D] = y[:, :D] + 0.8 * y[:, D:2 * D]       # positives near their anchors: both hinge branches occur
        yp = torch.tensor(y.astype(np.float64), requires_grad=True)
        loss = ns["triplet_loss"](None, yp.as_subclass(_Pred))
        loss.mean().backward()
        out[f"{name}/y"], out[f"{name}/loss"], out[f"{name}/grad"] = y, loss.detach().numpy(), yp.grad.numpy()
    a = rng.standard_normal((50, 64)).astype(np.float32)
    b = (a + 0.5 * rng.standard_normal((50, 64))).astype(np.float32)
    b[7] = a[7]                                              # zero distance: the epsilon clamp
    at, bt = torch.tensor(a.astype(np.float64), requires_grad=True), torch.tensor(b.astype(np.float64), requires_grad=True)
    d = ns["euclidean_distance"]((at, bt))
    yt = torch.tensor((np.arange(50) % 2).astype(np.float64)).reshape(-1, 1)
    cl = ns["contrastive_loss"](yt, d / 10.0)
    cl.backward()
    out["siamese/a"], out["siamese/b"], out["siamese/dist"] = a, b, d.detach().numpy()
    out["siamese/contrastive"], out["siamese/grad_a"], out["siamese/grad_b"] = cl.detach().numpy(), at.grad.numpy(), bt.grad.numpy()
    out["siamese/accuracy_default"] = ns["_accuracy"](yt, (d / 10.0).detach()).numpy()
    # networks/utils.py imports as is (numpy only); api.py:face_distance / compare_faces are cut out (the module builds a
    # Keras model at import time) and executed with the reference's own helpers in scope
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "deep_insight_face/networks/utils.py"))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    api_ns = {"np": np, "gaussian_kernel_dist_to_prob": ru.gaussian_kernel_dist_to_prob, "distance_to_proba": ru.distance_to_proba}
    for fn in ("face_distance", "compare_faces"):
        exec(function_source(os.path.join(REF, "deep_insight_face/api.py"), fn), api_ns)
    e = rng.standard_normal((6, 128)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    e[1] = e[0] + 0.02 * e[2]                                # a pair inside the 0.6 tolerance
    out["api/enc"] = e
    out["api/face_distance"] = np.array([api_ns["face_distance"](e[i], e[j]) for i in range(6) for j in range(6)])
    cf = [api_ns["compare_faces"]([e[i]], [e[j]]) for i in range(6) for j in range(6)]
    out["api/compare_faces"] = np.array(cf, dtype=np.float64)
    out["api/face_distance_empty_shape"] = np.array(api_ns["face_distance"]([], e[0]).shape)
    out["utils/distance"] = np.array([ru.distance(e[i], e[j]) for i in range(6) for j in range(6)], dtype=np.float64)
    dgrid = np.linspace(0.0, 3.0, 13)
    out["utils/dgrid"] = dgrid
    out["utils/distance_to_proba"] = ru.distance_to_proba(dgrid)
    out["utils/gaussian_kernel_1"] = ru.gaussian_kernel_dist_to_prob(dgrid)
    out["utils/gaussian_kernel_half"] = ru.gaussian_kernel_dist_to_prob(dgrid, 0.5)
    scores = rng.random(10)
    out["utils/scores"] = scores
    out["utils/calc_mean_score"] = np.float64(ru.calc_mean_score(scores))
    np.savez_compressed(os.path.join(HERE, "losses_reference.npz"), **out)
    print("wrote", len(out), "arrays to tests/golden/losses_reference.npz")


if __name__ == "__main__":
    main()
