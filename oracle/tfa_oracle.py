"""CPU restatement of tensorflow_addons' TripletHardLoss / TripletSemiHardLoss.  TEST INFRASTRUCTURE ONLY.

The reference calls `tfa.losses.TripletHardLoss()` and `tfa.losses.TripletSemiHardLoss()` with their defaults
(deep_insight_face/networks/triplet.py:196,209,211; sparse integer labels from training/triplet.py:72).  The
arithmetic lives in the third-party package tensorflow_addons (version UNPINNED: the reference's
requirements.txt is empty), files tensorflow_addons/losses/triplet.py and losses/metric_learning.py, which are
not under the reference tree and cannot be installed here.  The functions below restate the published algorithm
of those two files operation by operation (names kept), on the canonical fp32 pairwise matrix of
oracle/dif_oracle.c so mined indices compare bit for bit with the GPU; gradients come from torch autograd (fp64)
over the same literal forward, whose amax/amin split the cotangent over ties like tf.reduce_max / reduce_min.

PIN STATUS.  Semi-hard loss, pairwise_distance, _masked_maximum / _masked_minimum: PINNED against the one text of
this algorithm the reference holds - its copy of tf.contrib's triplet_semihard_loss, the function tensorflow_addons
ported (deep_insight_face/common/losses.py:151-308; executed by tests/golden/make_golden_semihard.py both as it is
and with its dangling `- 2.0 * matmul` line repaired, goldens in tests/golden/semihard_reference.npz, held by
tests/test_parity_cpu.py).  Hard loss: the reference holds no text of tfa's function, but with
distance_metric='squared-L2' it is anchor by anchor the reference's own BatchHardTripletLossEuclidean
(common/losses.py:54-85) whenever every anchor has a negative - PINNED in that mode against the goldens that class
produced (losses_reference.npz: loss to 1e-12, gradient to 1e-9 in fp64; same test file).  What stays PARITY
UNPINNED: the soft margin, and the step from these functions to tfa's defaults (sparse labels instead of one-hot,
the non-squared `distance_metric='L2'` the reference's call sites use - the sqrt / error-mask convention is pinned
through pairwise_distance above, its composition with the hard rule is taken from tfa's published source).
"""
from __future__ import annotations

import numpy as np

from . import c_oracle as orc

F32 = np.float32


def pairwise_distance(emb, squared=False):
    """metric_learning.pairwise_distance: ||a||^2 + ||b||^2 - 2ab, clamp at 0, sqrt with the error mask
    (entries <= 0 become exactly 0), diagonal forced to 0."""
    x = np.ascontiguousarray(emb, dtype=F32)
    sq = orc.row_sqnorm(x)
    d2 = (sq[:, None] + sq[None, :]) - F32(2.0) * orc.cross(x, x, 1)
    d2 = np.maximum(d2, F32(0.0))
    err = d2 <= 0
    d = d2 if squared else np.sqrt(d2 + err.astype(F32) * F32(1e-16))
    d = d * (~err).astype(F32)
    return (d * (F32(1.0) - np.eye(x.shape[0], dtype=F32))).astype(F32)


def _masked_minimum(data, mask):
    """triplet._masked_minimum (dim=1): min over the masked entries, computed as min((data - rowmax) * mask) + rowmax."""
    axis_max = data.max(axis=1, keepdims=True)
    return ((data - axis_max) * mask).min(axis=1, keepdims=True) + axis_max


def _masked_maximum(data, mask):
    axis_min = data.min(axis=1, keepdims=True)
    return ((data - axis_min) * mask).max(axis=1, keepdims=True) + axis_min


def triplet_hard(labels, emb, margin=1.0, soft=False, squared=False):
    """triplet.triplet_hard_loss: scalar mean over anchors; also the mined columns (first index on ties, -1 if
    the anchor has no other sample of its identity / no sample of another identity)."""
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    B = lab.shape[0]
    P = pairwise_distance(emb, squared)
    adjacency = lab[:, None] == lab[None, :]
    adjacency_not = (~adjacency).astype(F32)
    hard_negatives = _masked_minimum(P, adjacency_not)[:, 0]
    mask_positives = adjacency.astype(F32) - np.eye(B, dtype=F32)
    hard_positives = _masked_maximum(P, mask_positives)[:, 0]
    x = hard_positives - hard_negatives
    per = np.log1p(np.exp(x)) if soft else np.maximum(x + F32(margin), F32(0.0))
    pos_idx = np.full(B, -1, np.int32)
    neg_idx = np.full(B, -1, np.int32)
    rowmax = P.max(axis=1, keepdims=True)
    shifted = P - rowmax
    for b in range(B):
        pm = mask_positives[b] > 0
        if pm.any():
            pos_idx[b] = np.flatnonzero(pm & (P[b] == P[b][pm].max()))[0]
        nm = adjacency_not[b] > 0
        if nm.any():
            neg_idx[b] = np.flatnonzero(nm & (shifted[b] == shifted[b][nm].min()))[0]
    return {"loss": F32(per.astype(np.float64).mean()), "per_anchor": per.astype(F32), "pos_idx": pos_idx,
            "neg_idx": neg_idx, "hard_positives": hard_positives, "hard_negatives": hard_negatives}


def semihard_from_matrix(P, labels, margin=1.0):
    """The semi-hard rule of triplet.triplet_semihard_loss on a given pairwise matrix P (any float dtype), anchor
    by anchor (the [B*B, B] tiling of the original is the same arithmetic per (anchor b, positive a) pair):
    semi-hard negative = the closest negative farther than the positive (negatives_outside) or, when none is, the
    farthest negative (negatives_inside); sum of the hinges / number of positive pairs.  Pinned against the
    reference's copy of the tf.contrib ancestor of this function (deep_insight_face/common/losses.py:249-308,
    executed by tests/golden/make_golden_semihard.py) in tests/test_parity_cpu.py."""
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    B = lab.shape[0]
    T = P.dtype.type
    total = 0.0
    num_positives = 0
    for b in range(B):
        neg = lab != lab[b]
        pos = ~neg
        pos[b] = False
        num_positives += int(pos.sum())
        if not pos.any():
            continue
        row = P[b]
        rowmax = row.max()
        rowmin = row.min()
        inside = ((row - rowmin) * neg.astype(T)).max() + rowmin
        shifted = row - rowmax
        for a in np.flatnonzero(pos):
            mask = neg & (row > row[a])
            if mask.any():
                sh = (shifted * mask.astype(T)).min() + rowmax
            else:
                sh = inside
            lm = T(margin) + (row[a] - sh)
            total += float(max(lm, T(0.0)))
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = T(np.float64(total) / np.float64(num_positives)) if num_positives else T(np.nan)
    return {"loss": loss, "num_positives": num_positives}


def triplet_semihard(labels, emb, margin=1.0, squared=False):
    """triplet.triplet_semihard_loss: the semi-hard rule on the canonical fp32 pairwise matrix."""
    return semihard_from_matrix(pairwise_distance(emb, squared), labels, margin)


# ------------------------------------------------------------------ fp64 autograd shadow (torch CPU)
def _torch_pdist(x, squared):
    import torch

    sq = (x * x).sum(1, keepdim=True)
    d2 = (sq + sq.t()) - 2.0 * (x @ x.t())
    d2 = torch.clamp_min(d2, 0.0)
    err = (d2 <= 0).to(x.dtype)
    d = d2 if squared else torch.sqrt(d2 + err * 1e-16)
    d = d * (1.0 - err)
    return d * (1.0 - torch.eye(x.shape[0], dtype=x.dtype))


def _hinge(x):
    import torch

    return torch.where(x >= 0, x, torch.zeros_like(x))   # tf.maximum(x, 0): gradient to x where x >= 0


def _torch_loss(kind, P, lab, margin, soft):
    """The literal tfa forward from the pairwise matrix P on (any float dtype)."""
    import torch

    B = lab.shape[0]
    dt = P.dtype
    adj = lab[:, None] == lab[None, :]
    adj_not = (~adj).to(dt)
    eye = torch.eye(B, dtype=dt)

    def mmin(data, mask):
        ax = data.amax(1, keepdim=True)
        return ((data - ax) * mask).amin(1, keepdim=True) + ax

    def mmax(data, mask):
        ax = data.amin(1, keepdim=True)
        return ((data - ax) * mask).amax(1, keepdim=True) + ax

    if kind == "hard":
        hn = mmin(P, adj_not)
        hp = mmax(P, adj.to(dt) - eye)
        return (torch.log1p(torch.exp(hp - hn)) if soft else _hinge(hp - hn + margin)).mean()
    tile = P.repeat(B, 1)                                       # row (a, b) = P[b]
    mask = adj_not.repeat(B, 1) * (tile > P.t().reshape(-1, 1)).to(dt)
    mask_final = (mask.sum(1, keepdim=True) > 0).reshape(B, B).t()
    outside = mmin(tile, mask).reshape(B, B).t()
    inside = mmax(P, adj_not).repeat(1, B)
    semi = torch.where(mask_final, outside, inside)
    loss_mat = margin + (P - semi)
    mask_pos = adj.to(dt) - eye
    return _hinge(loss_mat * mask_pos).sum() / mask_pos.sum()


def torch_shadow_fp64(kind, labels, emb, margin=1.0, soft=False, squared=False):
    """fp64 end to end: autograd through the distance computation as well.  Selections are made on fp64
    distances, so it agrees with fp32 tfa only where no two fp32 distances tie after rounding."""
    import torch

    x = torch.tensor(np.asarray(emb, dtype=np.float64), requires_grad=True)
    lab = torch.tensor(np.asarray(labels).reshape(-1).astype(np.int64))
    loss = _torch_loss(kind, _torch_pdist(x, squared), lab, margin, soft)
    loss.backward()
    return float(loss.detach()), x.grad.numpy()


def torch_shadow(kind, labels, emb, margin=1.0, soft=False, squared=False):
    """Gradient oracle faithful to fp32 tfa: the literal forward runs in fp32 on the canonical fp32 matrix
    (so rounding-induced ties of (P - rowmax) split the cotangent exactly as tf.reduce_min does), autograd
    gives dL/dP, and the chain rule through pairwise_distance is applied in fp64:
    dP_ij/dx_i = (x_i - x_j) / P_ij (2 (x_i - x_j) squared), zero on the error-masked and diagonal entries.
    kind 'hard' | 'semihard' (the latter needs O(B^3) memory)."""
    import torch

    P0 = pairwise_distance(emb, squared)
    P = torch.tensor(P0, requires_grad=True)
    lab = torch.tensor(np.asarray(labels).reshape(-1).astype(np.int64))
    loss = _torch_loss(kind, P, lab, F32(margin).item(), soft)
    loss.backward()
    G = P.grad.numpy().astype(np.float64)
    x = np.asarray(emb, dtype=np.float64)
    P64 = P0.astype(np.float64)
    with np.errstate(divide="ignore"):
        fac = np.where(P64 > 0, 2.0 if squared else 1.0 / P64, 0.0)
    W = (G + G.T) * fac
    return float(loss.detach()), W.sum(1)[:, None] * x - W @ x


# ------------------------------------------------------------------ distance_metric="angular"
def _torch_angular(x):
    """metric_learning.angular_distance: l2_normalize the rows (x * rsqrt(max(sum x^2, 1e-12))),
    1 - x^ x^T, clamp at 0, zero diagonal."""
    import torch

    unit = x * torch.rsqrt(torch.clamp_min((x * x).sum(1, keepdim=True), 1e-12))
    a = torch.clamp_min(1.0 - unit @ unit.t(), 0.0)
    return a * (1.0 - torch.eye(x.shape[0], dtype=x.dtype))


def torch_shadow_angular_fp64(kind, labels, emb, margin=1.0, soft=False):
    """tfa's losses with distance_metric='angular', fp64 end to end (autograd through the normalisation too).
    The reference never passes this metric and holds no text of it: tfa's published source, PARITY UNPINNED."""
    import torch

    x = torch.tensor(np.asarray(emb, dtype=np.float64), requires_grad=True)
    lab = torch.tensor(np.asarray(labels).reshape(-1).astype(np.int64))
    loss = _torch_loss(kind, _torch_angular(x), lab, margin, soft)
    loss.backward()
    return float(loss.detach()), x.grad.numpy()

