"""BLAS-speed CPU statements of the loss / verification paths, for TIMING the reference-side cost.  TEST INFRASTRUCTURE ONLY.

The reference issues these computations as stock TensorFlow-CPU / numpy library calls (deep_insight_face/common/
losses.py:33-148 -> tf.matmul + tf.where + reduce_min/max + autodiff; evaluation/utility.py:36-171 -> numpy boolean
reductions in python threshold loops).  TensorFlow cannot be installed here, so bench.py's
`cpu_baseline` legs time the same op sequences on torch-CPU (multithreaded sgemm + autograd, same algorithmic cost
as TF's Eigen/oneDNN kernels) and numpy.  Values are NOT used for correctness - that is oracle/losses_oracle.py and
oracle/dif_oracle.c, which trade speed for a fixed summation order.  Only bench.py imports this module.
"""
from __future__ import annotations

import numpy as np


def _t(a, grad=False):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.requires_grad_(True) if grad else t


def batch_hard_step(labels, emb, alpha=0.35, cosine=True):
    """common/losses.py:33-51 (cosine) / :54-85 (squared L2): forward + backward of mean(loss), op for op."""
    import torch

    x = _t(emb, grad=True)
    lab = _t(np.asarray(labels, dtype=np.int64))
    pos = lab[:, None] == lab[None, :]
    if cosine:
        n = torch.nn.functional.normalize(x, dim=1, eps=1e-6)
        S = n @ n.T
        hp = torch.where(pos, S, torch.ones_like(S)).amin(dim=1)
        hn = torch.where(pos, -torch.ones_like(S), S).amax(dim=1)
        loss = torch.clamp(hn - hp + alpha, min=0)
    else:
        sq = (x * x).sum(1)
        Dm = sq[:, None] + sq[None, :] - 2.0 * (x @ x.T)
        hp = torch.where(pos, Dm, torch.zeros_like(Dm)).amax(dim=1)
        hn = torch.where(pos, Dm.max().expand_as(Dm), Dm).amin(dim=1)
        loss = torch.clamp(hp + alpha - hn, min=0)
    loss.mean().backward()
    return loss.detach().numpy(), x.grad.numpy()


def batch_all_step(labels, emb, alpha=0.35):
    """common/losses.py:131-148: forward + backward of mean(loss)."""
    import torch

    x = _t(emb, grad=True)
    lab = _t(np.asarray(labels, dtype=np.int64))
    pos = lab[:, None] == lab[None, :]
    n = torch.nn.functional.normalize(x, dim=1, eps=1e-6)
    S = n @ n.T
    one = torch.ones_like(S)
    pos_loss = (1.0 - torch.where(pos, S, one)).sum(1) / pos.sum(1)
    hp = torch.where(pos, S, one).amin(dim=1)
    valid = (~pos) & ((hp[:, None] - S) < alpha)
    neg_loss = torch.where(valid, S, torch.zeros_like(S)).sum(1) / (valid.sum(1) + 1.0)
    loss = pos_loss + neg_loss
    loss.mean().backward()
    return loss.detach().numpy(), x.grad.numpy()


def arcface_step(X, W, y, s=64.0, m=0.5):
    """ArcFace margin logits + softmax-CE (arXiv 1801.07698), fp32, forward + backward of mean(loss)."""
    import math

    import torch

    x = _t(X, grad=True)
    w = _t(W, grad=True)
    lab = _t(np.asarray(y, dtype=np.int64))
    cos = (torch.nn.functional.normalize(x, dim=1) @ torch.nn.functional.normalize(w, dim=1).T).clamp(-1, 1)
    cy = cos.gather(1, lab[:, None]).squeeze(1)
    sy = torch.sqrt(torch.clamp(1.0 - cy * cy, min=0))
    phi = cy * math.cos(m) - sy * math.sin(m)
    phi = torch.where(cy > math.cos(math.pi - m), phi, cy - m * math.sin(math.pi - m))
    logits = s * cos.scatter(1, lab[:, None], phi[:, None])
    loss = torch.nn.functional.cross_entropy(logits, lab, reduction="none")
    loss.mean().backward()
    return loss.detach().numpy(), x.grad.numpy(), w.grad.numpy()


def tfa_hard_step(labels, emb, margin=1.0):
    """tensorflow_addons TripletHardLoss defaults (networks/triplet.py:211): non-squared L2, hard margin, mean."""
    import torch

    x = _t(emb, grad=True)
    lab = _t(np.asarray(labels, dtype=np.int64))
    sq = (x * x).sum(1)
    P2 = torch.clamp(sq[:, None] + sq[None, :] - 2.0 * (x @ x.T), min=0)
    zero = P2 <= 0
    P = torch.sqrt(P2 + zero * 1e-16) * (~zero)
    P = P * (1.0 - torch.eye(P.shape[0]))
    adj = lab[:, None] == lab[None, :]
    hn = torch.where(~adj, P, P.max().expand_as(P)).amin(dim=1)
    hp = torch.where(adj & ~torch.eye(P.shape[0], dtype=torch.bool), P, torch.zeros_like(P)).amax(dim=1)
    loss = torch.clamp(hp - hn + margin, min=0).mean()
    loss.backward()
    return float(loss.detach()), x.grad.numpy()


def pair_distance(e1, e2, metric=0):
    """evaluation/utility.py:52-66 with the reference's own numpy calls."""
    if metric == 0:
        diff = np.subtract(e1, e2)
        return np.sum(np.square(diff), 1)
    dot = np.sum(np.multiply(e1, e2), axis=1)
    norm = np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1)
    return np.arccos(dot / norm) / np.pi


def roc_sweep(dist, issame, thresholds, n_folds=10):
    """The threshold loops of evaluation/utility.py:122-171 (calculate_roc) for precomputed distances: per fold a
    python loop over thresholds on the train split, then on the test split; four boolean reductions per call."""
    n = dist.shape[0]
    bounds = np.linspace(0, n, n_folds + 1).astype(int)
    tprs = np.zeros((n_folds, len(thresholds)))
    fprs = np.zeros((n_folds, len(thresholds)))
    acc = np.zeros(n_folds)

    def counts(thr, d, s):
        pred = np.less(d, thr)
        tp = np.sum(np.logical_and(pred, s))
        fp = np.sum(np.logical_and(pred, np.logical_not(s)))
        tn = np.sum(np.logical_and(np.logical_not(pred), np.logical_not(s)))
        fn = np.sum(np.logical_and(np.logical_not(pred), s))
        return tp, fp, tn, fn

    for f in range(n_folds):
        test = np.zeros(n, dtype=bool)
        test[bounds[f]:bounds[f + 1]] = True
        train = ~test
        acc_train = np.zeros(len(thresholds))
        for i, thr in enumerate(thresholds):
            tp, fp, tn, fn = counts(thr, dist[train], issame[train])
            acc_train[i] = (tp + tn) / max(1, train.sum())
        best = int(np.argmax(acc_train))
        for i, thr in enumerate(thresholds):
            tp, fp, tn, fn = counts(thr, dist[test], issame[test])
            tprs[f, i] = 0 if tp + fn == 0 else tp / (tp + fn)
            fprs[f, i] = 0 if fp + tn == 0 else fp / (fp + tn)
        tp, fp, tn, fn = counts(thresholds[best], dist[test], issame[test])
        acc[f] = (tp + tn) / max(1, test.sum())
    return tprs.mean(0), fprs.mean(0), acc


def num_threads() -> int:
    import torch

    return torch.get_num_threads()
