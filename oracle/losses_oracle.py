"""numpy restatement of the reference's triplet / siamese losses.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function follows the reference line by line (paths relative to the reference repository root) with the
B x B matrix computed in the canonical fp32 arithmetic of oracle/dif_oracle.c, so mined indices can be compared
bit for bit with the GPU.  TensorFlow semantics that the reference relies on but that live outside its tree [ext]:
  tf.nn.l2_normalize(x, 1) = x * rsqrt(max(sum(x^2), 1e-12));  tf.argmax = first maximum;
  reduce_min / reduce_max gradient = cotangent split evenly over every tied position;
  tf.where gradient flows only into the selected branch;  tf.maximum(x, 0) passes the gradient to x when x >= 0;
  Keras AUTO reduction = mean over the batch (default dloss = 1/B).
PINNED AGAINST THE REFERENCE'S OWN SOURCE (round 2): TensorFlow cannot be installed here and the reference ships
no golden vectors, so tests/golden/make_golden_losses.py IMPORTS deep_insight_face/common/
losses.py (and cuts triplet_loss / euclidean_distance / contrastive_loss / _accuracy out of networks/*.py) and runs
that code on a float64 torch stand-in for the two dozen TensorFlow / Keras-backend entry points it calls; the outputs
(losses, gradients of mean(loss) by autograd through the reference's own op sequence, AutoAlpha state) are committed
as tests/golden/losses_reference.npz and tests/test_parity_cpu.py holds these functions to them (2e-5).  What stays
an assumption is that the stand-in's one-liners mean what TensorFlow's ops mean (listed above).  ArcFace (absent from
the reference) remains parity-unpinned: its oracle is cross-checked against torch autograd only.
"""
from __future__ import annotations

import numpy as np

from . import c_oracle as orc

F32 = np.float32
EPS = F32(1e-12)


def _labels(labels) -> np.ndarray:
    labels = np.asarray(labels)
    if labels.ndim == 2:  # one-hot, common/losses.py:35 `tf.argmax(labels, axis=1)`
        labels = np.argmax(labels, axis=1)
    return labels.astype(np.int64)


def _first_index(mask_row: np.ndarray) -> int:
    nz = np.flatnonzero(mask_row)
    return int(nz[0]) if nz.size else -1


def cosine_matrix(emb):
    """common/losses.py:39-40: l2_normalize then n @ n.T, canonical fp32."""
    n = orc.normalize_rows(emb)
    return orc.cross(n, n, 1), n


def sqdist_matrix(emb):
    """common/losses.py:63-65: sq[:, None] + sq - 2 * (x @ x.T), canonical fp32, no clamp."""
    x = np.ascontiguousarray(emb, dtype=F32)
    sq = orc.row_sqnorm(x)
    ab = orc.cross(x, x, 1)
    return (sq[:, None] + sq[None, :]) - F32(2.0) * ab, sq


def _finish(loss, hp, hn, dists, pos_idx, neg_idx, grad, extra=None):
    out = {"loss": loss.astype(F32), "hardest_pos": hp, "hardest_neg": hn, "pos_idx": pos_idx.astype(np.int32),
           "neg_idx": neg_idx.astype(np.int32), "grad": grad,
           "stats": np.array([dists.astype(np.float64).mean(), hp.astype(np.float64).mean(),
                              hn.astype(np.float64).mean(), dists.max()], dtype=np.float64)}
    if extra:
        out.update(extra)
    return out


def _margin(basic, x, soft, B, dloss):
    """Hard margin of the reference (max(basic, 0), gradient where basic >= 0) or the soft margin
    log(1 + exp(x)) of arXiv 1703.07737 eq. 4 (extension, gradient sigmoid(x))."""
    dl = (1.0 / B) if dloss is None else np.asarray(dloss, dtype=np.float64)
    if not soft:
        return np.maximum(basic, F32(0.0)), np.where(basic >= 0, dl, 0.0)
    x64 = x.astype(np.float64)
    return np.logaddexp(0.0, x64).astype(F32), dl / (1.0 + np.exp(-x64))


def batch_hard_cosine(labels, emb, alpha=0.35, dloss=None, soft=False):
    """common/losses.py:33-51 BatchHardTripletLoss."""
    lab = _labels(labels)
    x = np.ascontiguousarray(emb, dtype=F32)
    B = x.shape[0]
    S, n = cosine_matrix(x)
    pos = lab[:, None] == lab[None, :]                                  # :38
    posv = np.where(pos, S, F32(1.0))                                    # :42
    hp = posv.min(axis=1)                                                # :43
    negv = np.where(pos, F32(-1.0), S)                                   # :45
    hn = negv.max(axis=1)                                                # :46
    basic = (hn - hp) + F32(alpha)                                       # :47
    loss, g = _margin(basic, hn - hp, soft, B, dloss)                    # :51
    pos_idx = np.full(B, -1, dtype=np.int64)
    neg_idx = np.full(B, -1, dtype=np.int64)
    for i in range(B):
        pos_idx[i] = _first_index(pos[i] & (S[i] == hp[i]))             # -1 when the filler (1.0) wins strictly
        neg_idx[i] = _first_index(~pos[i] & (S[i] == hn[i]))
    tied_p = posv == hp[:, None]
    tied_n = negv == hn[:, None]
    G = np.zeros((B, B), dtype=np.float64)
    G += (g / tied_n.sum(1))[:, None] * (tied_n & ~pos)
    G -= (g / tied_p.sum(1))[:, None] * (tied_p & pos)
    n64 = n.astype(np.float64)
    dN = (G + G.T) @ n64
    ss = (x.astype(np.float64) ** 2).sum(1)
    inv = 1.0 / np.sqrt(np.maximum(ss, float(EPS)))
    proj = dN - n64 * (n64 * dN).sum(1, keepdims=True)
    grad = np.where((ss < float(EPS))[:, None], dN, proj) * inv[:, None]
    return _finish(loss, hp, hn, S, pos_idx, neg_idx, grad.astype(np.float64))


def batch_hard_euclidean(labels, emb, alpha=0.35, dloss=None, soft=False):
    """common/losses.py:54-85 BatchHardTripletLossEuclidean (and :88-128 with alpha = auto_alpha)."""
    lab = _labels(labels)
    x = np.ascontiguousarray(emb, dtype=F32)
    B = x.shape[0]
    Dm, _ = sqdist_matrix(x)
    pos = lab[:, None] == lab[None, :]                                  # :60
    posv = np.where(pos, Dm, F32(0.0))                                   # :67
    hp = posv.max(axis=1)                                                # :68
    gmax = Dm.max()                                                      # :70 tf.reduce_max(dists)
    negv = np.where(pos, gmax, Dm)                                       # :70
    hn = negv.min(axis=1)                                                # :71
    basic = (hp + F32(alpha)) - hn                                       # :81
    loss, g = _margin(basic, hp - hn, soft, B, dloss)                    # :85
    pos_idx = np.full(B, -1, dtype=np.int64)
    neg_idx = np.full(B, -1, dtype=np.int64)
    for i in range(B):
        pos_idx[i] = _first_index(pos[i] & (Dm[i] == hp[i]))            # -1 when the filler (0) wins strictly
        neg_idx[i] = _first_index(~pos[i] & (Dm[i] == hn[i]))           # -1 when only the max(dists) filler is left
    tied_p = posv == hp[:, None]
    tied_n = negv == hn[:, None]
    G = np.zeros((B, B), dtype=np.float64)
    G += (g / tied_p.sum(1))[:, None] * (tied_p & pos)
    G -= (g / tied_n.sum(1))[:, None] * (tied_n & ~pos)
    # cotangent of the max(dists) fillers (tied filler positions are the positive columns)
    share = -(g / tied_n.sum(1)) * (tied_n & pos).sum(1)
    at_max = Dm == gmax
    G += share.sum() / at_max.sum() * at_max
    W = G + G.T
    x64 = x.astype(np.float64)
    grad = 2.0 * (W.sum(1, keepdims=True) * x64 - W @ x64)
    return _finish(loss, hp, hn, Dm, pos_idx, neg_idx, grad)


def batch_all_cosine(labels, emb, alpha=0.35, dloss=None):
    """common/losses.py:131-148 BatchAllTripletLoss.  The `valid` mask is piecewise constant, so the gradient is
    -1/n_pos on the positive columns and +1/(n_valid + 1) on the valid negative columns (tf.where + reduce_sum)."""
    lab = _labels(labels)
    x = np.ascontiguousarray(emb, dtype=F32)
    B = x.shape[0]
    S, n = cosine_matrix(x)
    pos = lab[:, None] == lab[None, :]
    posv = np.where(pos, S, F32(1.0))
    pos_loss = (F32(1.0) - posv).sum(1) / pos.sum(1).astype(F32)        # :140-141
    hp = posv.min(1, keepdims=True)                                      # :142
    valid = ~pos & ((hp - S) < F32(alpha))                               # :144
    neg_loss = np.where(valid, S, F32(0.0)).sum(1) / (valid.sum(1).astype(F32) + F32(1.0))  # :145-147
    g = (np.full(B, 1.0 / B) if dloss is None else np.asarray(dloss, dtype=np.float64))[:, None]
    G = -g / pos.sum(1, keepdims=True) * pos + g / (valid.sum(1, keepdims=True) + 1.0) * valid
    n64 = n.astype(np.float64)
    dN = (G + G.T) @ n64
    ss = (x.astype(np.float64) ** 2).sum(1)
    inv = 1.0 / np.sqrt(np.maximum(ss, float(EPS)))
    proj = dN - n64 * (n64 * dN).sum(1, keepdims=True)
    grad = np.where((ss < float(EPS))[:, None], dN, proj) * inv[:, None]
    return {"loss": (pos_loss + neg_loss).astype(F32), "pos_loss": pos_loss, "neg_loss": neg_loss,
            "valid_count": valid.sum(1), "grad": grad}


def triplet_apn(y_pred, alpha=0.4, dloss=None):
    """networks/triplet.py:16-46 triplet_loss on [B, 3D] rows (anchor | positive | negative)."""
    y = np.ascontiguousarray(y_pred, dtype=F32)
    B, D3 = y.shape
    D = D3 // 3
    a, p, n = y[:, :D], y[:, D:2 * D], y[:, 2 * D:3 * D]
    dp = np.array([orc.lib().dif_or_canon_sqdist(np.ascontiguousarray(a[i]).ctypes.data,
                                                 np.ascontiguousarray(p[i]).ctypes.data, D) for i in range(B)], dtype=F32)
    dn = np.array([orc.lib().dif_or_canon_sqdist(np.ascontiguousarray(a[i]).ctypes.data,
                                                 np.ascontiguousarray(n[i]).ctypes.data, D) for i in range(B)], dtype=F32)
    basic = (dp - dn) + F32(alpha)                                       # :43
    loss = np.maximum(basic, F32(0.0))                                   # :44
    g = np.where(basic >= 0, 1.0 if dloss is None else np.asarray(dloss, dtype=np.float64), 0.0)[:, None]
    a64, p64, n64 = a.astype(np.float64), p.astype(np.float64), n.astype(np.float64)
    grad = np.concatenate([2 * g * (n64 - p64), -2 * g * (a64 - p64), 2 * g * (a64 - n64)], axis=1)
    return {"loss": loss, "grad": grad}


def euclidean_distance(x, y, eps=1e-7):
    """networks/siamese.py:22-24: sqrt(max(sum((x - y)^2), K.epsilon())), keepdims."""
    d = orc.pair_distance(x, y, 0)
    return np.sqrt(np.maximum(d, F32(eps))).astype(F32)[:, None]


def contrastive_loss(y_true, dist, margin=1.0):
    """networks/siamese.py:32-39: mean(y * d^2 + (1 - y) * max(margin - d, 0)^2)."""
    y = np.asarray(y_true, dtype=np.float64).reshape(-1)
    d = np.asarray(dist, dtype=np.float64).reshape(-1)
    m = np.maximum(margin - d, 0.0)
    return float(np.mean(y * d * d + (1.0 - y) * m * m)), (2 * y * d - 2 * (1 - y) * m) / y.size


def siamese_accuracy(y_true, dist, threshold=0.4):
    """networks/siamese.py:42-45: mean(y == (d < threshold)), reference default threshold 0.4."""
    y = np.asarray(y_true).reshape(-1)
    d = np.asarray(dist).reshape(-1)
    return float(np.mean(y == (d < threshold).astype(y.dtype)))


def arcface(X, W, y, s=64.0, m=0.5):
    """ArcFace additive-angular-margin logits + softmax cross-entropy (arXiv 1801.07698), fp64.

    ABSENT from the reference: the specification is this build's:
      xh = l2_normalize(x), wh = l2_normalize(w), cos = clip(xh . wh, -1, 1);
      target logit  s * cos(theta + m)  if theta + m <= pi  else  s * (cos - m * sin(m))  ("easy margin" fallback
      of the reference implementation of the paper: cos(theta) - mm with mm = sin(pi - m) * m when cos <= cos(pi - m));
      other logits s * cos;  loss_i = -log softmax(logits_i)[y_i];  gradients of mean(loss) wrt X and W.
    """
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    y = np.asarray(y, dtype=np.int64)
    B = X.shape[0]
    xs = (X ** 2).sum(1)
    ws = (W ** 2).sum(1)
    xi = 1.0 / np.sqrt(np.maximum(xs, 1e-12))
    wi = 1.0 / np.sqrt(np.maximum(ws, 1e-12))
    xh, wh = X * xi[:, None], W * wi[:, None]
    cos = np.clip(xh @ wh.T, -1.0, 1.0)
    ct = cos[np.arange(B), y]
    sin_t = np.sqrt(np.maximum(1.0 - ct * ct, 0.0))
    cos_m, sin_m = np.cos(m), np.sin(m)
    th = np.cos(np.pi - m)
    mm = np.sin(np.pi - m) * m
    phi = np.where(ct > th, ct * cos_m - sin_t * sin_m, ct - mm)
    # d phi / d cos
    with np.errstate(divide="ignore", invalid="ignore"):
        dphi = np.where(ct > th, cos_m + np.where(sin_t > 0, ct / np.maximum(sin_t, 1e-300), 0.0) * sin_m, 1.0)
    logits = s * cos
    logits[np.arange(B), y] = s * phi
    mx = logits.max(1, keepdims=True)
    ex = np.exp(logits - mx)
    Z = ex.sum(1, keepdims=True)
    p = ex / Z
    loss = -(logits[np.arange(B), y] - mx[:, 0] - np.log(Z[:, 0]))
    dlog = p.copy()
    dlog[np.arange(B), y] -= 1.0
    dlog /= B
    dcos = s * dlog
    dcos[np.arange(B), y] *= dphi
    # clip passes gradient only strictly inside (-1, 1) bounds (tf.clip_by_value)
    raw = xh @ wh.T
    dcos = np.where((raw < -1.0) | (raw > 1.0), 0.0, dcos)
    dxh = dcos @ wh
    dwh = dcos.T @ xh
    dX = xi[:, None] * (dxh - xh * (xh * dxh).sum(1, keepdims=True))
    dW = wi[:, None] * (dwh - wh * (wh * dwh).sum(1, keepdims=True))
    return {"loss": loss, "dX": dX, "dW": dW, "cos_target": ct}
