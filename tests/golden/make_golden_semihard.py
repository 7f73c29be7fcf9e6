"""Generates tests/golden/semihard_reference.npz by EXECUTING THE REFERENCE'S OWN SEMI-HARD TEXT, in the build container:

    python tests/golden/make_golden_semihard.py

The two tensorflow_addons losses the reference trains with (a8: networks/triplet.py:196,209,211) are third-party code
that is not under /root/reference.  Their ancestor is: deep_insight_face/common/losses.py:151-308 carries a copy of
tf.contrib.losses.metric_learning.triplet_semihard_loss (`pairwise_distance`, `masked_maximum`, `masked_minimum`,
`triplet_loss_adapted_from_tf`), the function tensorflow_addons ported operation by operation into
TripletSemiHardLoss / metric_learning.pairwise_distance.  The reference never calls it (a6, dead code: the import at
networks/triplet.py:197 is commented out), but it is the only text of that algorithm the reference holds, so it is
what pins oracle/tfa_oracle.py.  common/losses.py is IMPORTED unmodified on the float64 torch stand-in of
make_golden_losses.py, extended with the `math_ops` / `array_ops` entry points this code calls, and run three ways:

 * masked_maximum / masked_minimum as they are, on random data and masks (rows with an empty mask included);
 * triplet_loss_adapted_from_tf AS IT IS.  Its pairwise_distance drops the `- 2.0 * matmul(...)` term (the
   expression sits alone on line 183 after the closed math_ops.add(...)), so what it computes is the semi-hard rule
   on the matrix |a|^2 + |b|^2: a pin of the selection logic, the hinge and the normalisation, not of the distances;
 * the same module with ONE repair, made on the source text before it is compiled and asserted to match exactly
   once: the dangling line is joined to the statement above it (`keepdims=True)) \\` + newline + `- 2.0 * ...`).
   That is tf.contrib's function as published: pairwise_distance (squared and not) and the squared-L2 semi-hard
   loss with margin 1, with gradients (autograd of the stand-in, fp64, d loss / d embeddings).

Nothing here is product code and nothing of a6 is reproduced in the product (known-divergent, dead);
/root/reference does not exist on the GPU box, only the .npz travels.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

import make_golden_losses as G

REF = G.REF
HERE = os.path.dirname(os.path.abspath(__file__))
DT = G.DT
LOSSES_PY = os.path.join(REF, "deep_insight_face/common/losses.py")
DANGLING = "keepdims=True))\n    - 2.0 * math_ops.matmul("
JOINED = "keepdims=True)) \\\n    - 2.0 * math_ops.matmul("


def _axis(a):
    return a[0] if isinstance(a, (list, tuple)) and len(a) == 1 else a


def _reduce(fn):
    def f(x, axis=None, keepdims=False, name=None):
        return fn(x) if axis is None else fn(x, dim=_axis(axis), keepdim=keepdims)
    return f


def fill_ops(tf, mods):
    """math_ops / array_ops as lines 151-308 use them.  reduce_min / reduce_max split the cotangent over ties
    (torch.amin / amax), maximum sends it to x where x >= y - TensorFlow's rules, as in make_golden_losses.py."""
    m = mods["tensorflow.python.ops"].math_ops
    a = mods["tensorflow.python.ops"].array_ops
    m.add = lambda x, y, name=None: torch.as_tensor(x, dtype=DT) + y
    m.multiply = lambda x, y: x * y
    m.truediv = lambda x, y, name=None: x / y
    m.square = lambda x: x * x
    m.sqrt = torch.sqrt
    m.matmul = lambda x, y: x @ y
    m.maximum = G._maximum
    m.less_equal = lambda x, y: x <= y
    m.greater = lambda x, y: x > y
    m.equal = lambda x, y: x == y
    m.logical_not = lambda x: ~x
    m.logical_and = lambda x, y: x & y
    m.to_float = lambda x: x.to(DT)
    m.cast = lambda x, dtype=None: x.to(dtype)
    m.reduce_sum = _reduce(torch.sum)
    m.reduce_min = _reduce(torch.amin)
    m.reduce_max = _reduce(torch.amax)
    a.transpose = lambda x: x.t()
    a.shape = lambda x: tuple(x.shape)
    a.size = lambda x: int(x.numel())
    a.ones_like = torch.ones_like
    a.ones = lambda shape: torch.ones(tuple(int(s) for s in shape), dtype=DT)
    a.diag = torch.diag
    a.reshape = lambda x, shape: torch.reshape(x, tuple(int(s) for s in shape))
    a.tile = lambda x, reps: x.repeat(*[int(r) for r in reps])
    a.where = lambda c, x, y: torch.where(c, x, y)


def load_reference(repaired):
    tf, mods = G.make_tf()
    fill_ops(tf, mods)
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        if not repaired:
            spec = importlib.util.spec_from_file_location("ref_losses_a6", LOSSES_PY)
            ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref)                     # the reference's module, unmodified
        else:
            src = open(LOSSES_PY).read()
            assert src.count(DANGLING) == 1, "the dangling `- 2.0 * matmul` line of pairwise_distance was not found"
            ref = types.ModuleType("ref_losses_a6_repaired")
            exec(compile(src.replace(DANGLING, JOINED), LOSSES_PY, "exec"), ref.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


CASES = [  # name, P, K, D, noise, seed, flags      (pk_batch of make_golden_losses.py; regenerated by the tests)
    ("c1", 18, 4, 128, 1.0, 20, {}),
    ("c1_tight", 18, 4, 128, 0.3, 21, {}),
    ("wide", 33, 3, 100, 1.5, 22, {}),
    ("ragged", 11, 5, 30, 0.6, 23, {}),
    ("dups", 8, 4, 64, 0.7, 24, {"duplicates": True}),
    ("singletons_mixed", 24, 1, 64, 1.0, 25, {}),           # case_inputs(): two identities are merged so positives exist
    ("raw", 12, 4, 32, 1.0, 26, {}),                        # unscaled: squared distances of ~100, few active hinges
]


def case_inputs(name, P, K, D, noise, seed, flags):
    """Embeddings scaled so that squared distances between identities are ~2 (margin 1: active and inactive hinges,
    negatives inside and outside the positive distance all occur); 'raw' keeps the N(0, 1) scale."""
    emb, lab = G.pk_batch(P, K, D, noise, seed, **flags)
    if name == "singletons_mixed":
        lab = lab.copy()
        lab[lab == 1] = 0                                    # one identity of two samples, 22 singletons
        lab[lab > 1] -= 1
    if name != "raw":
        emb = (emb * np.float32(np.sqrt(2.0 / (2.0 * D * (noise * noise + 1.0))))).astype(np.float32)
    return emb, lab


def run(fn, emb, lab):
    x = torch.tensor(emb.astype(np.float64), requires_grad=True)
    onehot = torch.tensor(np.eye(int(lab.max()) + 1)[lab])
    loss = fn(onehot, x)                                     # the reference's code
    loss.backward()
    return np.float64(loss.item()), x.grad.numpy()


def main():
    ref, fixed = load_reference(False), load_reference(True)
    out = {}
    rng = np.random.default_rng(31)
    data = rng.standard_normal((40, 23))
    data[5] = np.abs(data[5])                                # a row of positive entries: the fill value 0 is below them
    data[6] = -np.abs(data[6])                               # and a row of negative ones
    mask = rng.random((40, 23)) < 0.4
    mask[7] = False                                          # empty mask: the result is the row extreme
    mask[8] = True
    out["masked/data"], out["masked/mask"] = data, mask
    out["masked/maximum"] = ref.masked_maximum(torch.tensor(data), torch.tensor(mask)).numpy()
    out["masked/minimum"] = ref.masked_minimum(torch.tensor(data), torch.tensor(mask)).numpy()
    for case in CASES:
        name = case[0]
        emb, lab = case_inputs(*case)
        x = torch.tensor(emb.astype(np.float64))
        out[f"{name}/as_is/pdist_squared"] = ref.pairwise_distance(x, squared=True).numpy()
        out[f"{name}/as_is/loss"], out[f"{name}/as_is/grad"] = run(ref.triplet_loss_adapted_from_tf, emb, lab)
        out[f"{name}/repaired/pdist_squared"] = fixed.pairwise_distance(x, squared=True).numpy()
        out[f"{name}/repaired/pdist"] = fixed.pairwise_distance(x, squared=False).numpy()
        out[f"{name}/repaired/loss"], out[f"{name}/repaired/grad"] = run(fixed.triplet_loss_adapted_from_tf, emb, lab)
    np.savez_compressed(os.path.join(HERE, "semihard_reference.npz"), **out)
    print("wrote", len(out), "arrays to tests/golden/semihard_reference.npz")
    for case in CASES:
        print(f"  {case[0]:18s} as is {out[case[0] + '/as_is/loss']:.6f}   repaired {out[case[0] + '/repaired/loss']:.6f}")


if __name__ == "__main__":
    main()
