"""Drop-in for deep_insight_face/evaluation/utility.py:10-188 (pair distance + threshold sweeps) on B200 kernels.

Same function names, argument order, return values and error behaviour as the reference.  The python
threshold loops of the reference ((400 + 400) x 10 + 4000 x 10 calls of four boolean reductions,
utility.py:106-107,154-166) become ONE device pass per distance vector: csrc/pairs.cu histograms every
pair against all thresholds for all folds at once and the integer counts are turned into the reference's
float64 ratios here, formula for formula - so tpr / fpr / accuracy / f1 / val / far are bit-identical to
the reference whenever the distances are.
"""
from __future__ import annotations

import math
import os

import numpy as np

from .. import _ffi


def _torch():
    import torch

    return torch


def _dev(a, dtype):
    torch = _torch()
    if _ffi.is_device_tensor(a):
        return a.contiguous().to(dtype)
    return torch.from_numpy(np.ascontiguousarray(_ffi.host_array(a, None))).to("cuda", dtype=dtype).contiguous()


def kfold_ids(n: int, n_splits: int) -> np.ndarray:
    """Fold of each index under sklearn KFold(n_splits, shuffle=False): contiguous folds, the first n % n_splits
    one element longer (utility.py:90,134)."""
    sizes = np.full(n_splits, n // n_splits, dtype=np.int64)
    sizes[: n % n_splits] += 1
    return np.repeat(np.arange(n_splits, dtype=np.int32), sizes)


def distance(embeddings1, embeddings2, distance_metric=0, _mean=None):
    """utility.py:52-66.  metric 0: sum((a-b)^2); metric 1: arccos(cosine)/pi; else RuntimeError."""
    if distance_metric not in (0, 1):
        raise RuntimeError('Undefined distance metric %d' % distance_metric)
    lib = _ffi.load_library()
    _ffi.init()
    if _ffi.is_device_tensor(embeddings1):
        torch = _torch()
        e1, e2 = _dev(embeddings1, torch.float32), _dev(embeddings2, torch.float32)
        out = torch.empty(e1.shape[0], dtype=torch.float32, device=e1.device)
        _ffi.check(lib.dif_pair_distance(_ffi.ptr(e1), _ffi.ptr(e2), e1.shape[0], e1.shape[1], distance_metric,
                                         _ffi.ptr(_mean), _ffi.ptr(out), _ffi.current_stream_ptr(e1.device)))
        return out
    e1 = _ffi.host_array(embeddings1, np.float32)
    e2 = _ffi.host_array(embeddings2, np.float32, e1.shape)
    out = np.empty(e1.shape[0], dtype=np.float32)
    mean = None if _mean is None else _ffi.host_array(_mean, np.float32, (e1.shape[1],))
    _ffi.check(lib.dif_pair_distance_host(_ffi.ptr(e1), _ffi.ptr(e2), e1.shape[0], e1.shape[1], distance_metric,
                                          _ffi.ptr(mean), _ffi.ptr(out)))
    return out


def get_emd_distance(embeddings1, embeddings2, distance_metric=0):
    """utility.py:174-188, the single-pair twin: two 1-D embeddings and metric 0 give a scalar, which is the form the
    reference calls it in (evals.py:119).  For 2-D inputs the reference's metric 0 sums over axis 0 (one number per
    embedding dimension, summed over the pairs - an artefact of the 1-D twin); here 2-D inputs give the per-pair
    distances of `distance`, which is also what the reference's metric 1 branch returns."""
    e1 = np.atleast_2d(_ffi.host_array(embeddings1, np.float32))
    e2 = np.atleast_2d(_ffi.host_array(embeddings2, np.float32))
    d = distance(e1, e2, distance_metric)
    return d[0] if np.ndim(embeddings1) == 1 and distance_metric == 0 else d


def threshold_counts(dist, actual_issame, thresholds, fold=None, n_folds=1) -> np.ndarray:
    """counts[f, t] = (tp, fp, tn, fn) over the pairs of fold f with predict = dist < thresholds[t] (np.less)."""
    lib = _ffi.load_library()
    _ffi.init()
    thr = np.ascontiguousarray(np.atleast_1d(np.asarray(thresholds, dtype=np.float64)))
    same = np.ascontiguousarray(np.asarray(actual_issame).astype(bool).astype(np.uint8))
    d = _ffi.host_array(dist, np.float32)
    n = min(d.shape[0], same.shape[0])
    fo = None if fold is None else np.ascontiguousarray(fold, dtype=np.int32)
    counts = np.empty((n_folds, thr.shape[0], 4), dtype=np.int64)
    _ffi.check(lib.dif_threshold_sweep_host(_ffi.ptr(d), _ffi.ptr(same), _ffi.ptr(fo), n, n_folds, _ffi.ptr(thr),
                                            thr.shape[0], _ffi.ptr(counts)))
    return counts


def _acc_from_counts(tp, fp, tn, fn, size):
    """utility.py:43-49, on integer counts."""
    tp, fp, tn, fn = int(tp), int(fp), int(tn), int(fn)
    tpr = 0 if (tp + fn == 0) else float(tp) / float(tp + fn)
    fpr = 0 if (fp + tn == 0) else float(fp) / float(fp + tn)
    acc = float(tp + tn) / size
    precision = 0 if (tp + fp == 0) else float(tp) / float(tp + fp)
    recall = 0 if (tp + fn == 0) else float(tp) / float(tp + fn)
    f1score = 0 if float(precision + recall) == 0.0 else 2 * (float(precision * recall) / float(precision + recall))
    return tpr, fpr, acc, f1score


def calculate_accuracy(threshold, dist, actual_issame, display_cm=False):
    """utility.py:36-49."""
    c = threshold_counts(dist, actual_issame, [threshold])[0, 0]
    return _acc_from_counts(*c, np.size(dist))


def _rates_from_counts(c):
    """Vectorised utility.py:43-45,73-77 over a [T, 4] count block: (tpr, fpr, acc) as float64 arrays.  The same
    IEEE divisions as the scalar formulas, so the values are bit-identical to the reference's python floats."""
    c = np.asarray(c, dtype=np.float64)
    tp, fp, tn, fn = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        tpr = np.where(tp + fn == 0, 0.0, tp / (tp + fn))
        fpr = np.where(fp + tn == 0, 0.0, fp / (fp + tn))
        acc = (tp + tn) / (tp + fp + tn + fn)
    return tpr, fpr, acc


def _val_far_from_counts(tp, fp, tn, fn):
    """utility.py:73-77."""
    n_same, n_diff = int(tp + fn), int(fp + tn)
    val = 0 if n_same == 0 else float(tp) / float(n_same)
    far = 0 if n_diff == 0 else float(fp) / float(n_diff)
    return val, far


def calculate_val_far(threshold, dist, actual_issame):
    """utility.py:69-77."""
    return _val_far_from_counts(*threshold_counts(dist, actual_issame, [threshold])[0, 0])


def _fold_distances(embeddings1, embeddings2, fold, nrof_folds, distance_metric, subtract_mean):
    """dist per fold: one vector shared by all folds, or (subtract_mean) one per fold with the mean of the
    fold's TRAIN rows removed first (utility.py:98-102,144-148)."""
    e1 = _ffi.host_array(embeddings1, np.float32)
    e2 = _ffi.host_array(embeddings2, np.float32)
    if not subtract_mean:
        d = distance(e1, e2, distance_metric)
        return [d] * nrof_folds
    torch = _torch()
    lib = _ffi.load_library()
    _ffi.init()
    n, D = e1.shape
    d1, d2 = _dev(e1, torch.float32), _dev(e2, torch.float32)
    begin = np.concatenate([[0], np.cumsum(np.bincount(fold, minlength=nrof_folds))]).astype(np.int64)
    fb = torch.from_numpy(begin).cuda()
    ws = torch.empty(nrof_folds * D, dtype=torch.float64, device="cuda")
    means = torch.empty((nrof_folds, D), dtype=torch.float32, device="cuda")
    st = _ffi.current_stream_ptr(d1.device)
    _ffi.check(lib.dif_fold_mean(_ffi.ptr(d1), _ffi.ptr(d2), _ffi.ptr(fb), nrof_folds, D, _ffi.ptr(ws), _ffi.ptr(means), st))
    return [distance(d1, d2, distance_metric, _mean=means[f]).cpu().numpy() for f in range(nrof_folds)]


def calculate_roc(thresholds, embeddings1, embeddings2, actual_issame, nrof_folds=10, distance_metric=0,
                  subtract_mean=False):
    """utility.py:122-171."""
    assert embeddings1.shape[0] == embeddings2.shape[0]
    assert embeddings1.shape[1] == embeddings2.shape[1]
    actual_issame = np.asarray(actual_issame)
    nrof_pairs = min(len(actual_issame), embeddings1.shape[0])
    fold = kfold_ids(nrof_pairs, nrof_folds)
    dists = _fold_distances(embeddings1[:nrof_pairs], embeddings2[:nrof_pairs], fold, nrof_folds, distance_metric,
                            subtract_mean)
    return roc_from_distances(thresholds, dists, actual_issame[:nrof_pairs], nrof_folds, per_fold=subtract_mean)


def roc_from_distances(thresholds, dists, actual_issame, nrof_folds=10, per_fold=False):
    """The k-fold threshold selection of utility.py:150-171 on PRECOMPUTED distances: `dists` is one [N] vector, or
    (per_fold=True, the subtract_mean case) a list with one vector per fold.  Counts come from one histogram pass
    per vector; the ratios are the reference's float64 formulas on integer counts, so the result is bit-identical
    to the reference's whenever the distances are (tests feed it the reference's own distances)."""
    actual_issame = np.asarray(actual_issame)
    nrof_pairs = len(actual_issame)
    nrof_thresholds = len(thresholds)
    fold = kfold_ids(nrof_pairs, nrof_folds)
    if not per_fold and not isinstance(dists, (list, tuple)):
        dists = [dists] * nrof_folds
    tprs = np.zeros((nrof_folds, nrof_thresholds))
    fprs = np.zeros((nrof_folds, nrof_thresholds))
    accuracy = np.zeros((nrof_folds))
    f1scores = np.zeros((nrof_folds))
    shared = None
    for fold_idx in range(nrof_folds):
        if per_fold or shared is None:
            shared = threshold_counts(dists[fold_idx], actual_issame, thresholds, fold, nrof_folds)
        counts = shared
        test = counts[fold_idx]                       # [T, 4]
        train = counts.sum(axis=0) - test             # exact integer arithmetic
        n_test = int(test[0].sum())
        _, _, acc_train = _rates_from_counts(train)
        best_threshold_index = np.argmax(acc_train)   # utility.py:159, first maximum
        tprs[fold_idx], fprs[fold_idx], _ = _rates_from_counts(test)
        _, _, accuracy[fold_idx], f1scores[fold_idx] = _acc_from_counts(*test[best_threshold_index], n_test)
    tpr = np.mean(tprs, 0)
    fpr = np.mean(fprs, 0)
    return tpr, fpr, accuracy, f1scores


def _interp_threshold(far_train, thresholds, far_target):
    """utility.py:109 `interpolate.interp1d(far_train, thresholds, kind='slinear')(far_target)`.  far_train is
    non-decreasing with plateaus, which scipy >= 1.12 rejects (the reference crashes here);
    this is the same piecewise-linear inverse with duplicates allowed."""
    return float(np.interp(far_target, far_train, thresholds))


def calculate_val(thresholds, embeddings1, embeddings2, actual_issame, far_target, nrof_folds=10, distance_metric=0,
                  subtract_mean=False):
    """utility.py:80-119."""
    assert embeddings1.shape[0] == embeddings2.shape[0]
    assert embeddings1.shape[1] == embeddings2.shape[1]
    actual_issame = np.asarray(actual_issame)
    nrof_pairs = min(len(actual_issame), embeddings1.shape[0])
    nrof_thresholds = len(thresholds)
    fold = kfold_ids(nrof_pairs, nrof_folds)
    val = np.zeros(nrof_folds)
    far = np.zeros(nrof_folds)
    dists = _fold_distances(embeddings1[:nrof_pairs], embeddings2[:nrof_pairs], fold, nrof_folds, distance_metric,
                            subtract_mean)
    shared = None
    for fold_idx in range(nrof_folds):
        if subtract_mean or shared is None:
            shared = threshold_counts(dists[fold_idx], actual_issame[:nrof_pairs], thresholds, fold, nrof_folds)
        train = shared.sum(axis=0) - shared[fold_idx]
        _, far_train, _ = _rates_from_counts(train)   # far = fp / n_diff, utility.py:76-77
        if np.max(far_train) >= far_target:
            threshold = _interp_threshold(far_train, thresholds, far_target)
        else:
            threshold = 0.0
        c = threshold_counts(dists[fold_idx], actual_issame[:nrof_pairs], [threshold], fold, nrof_folds)[fold_idx, 0]
        val[fold_idx], far[fold_idx] = _val_far_from_counts(*c)
    val_mean = np.mean(val)
    far_mean = np.mean(far)
    val_std = np.std(val)
    return val_mean, val_std, far_mean


def evaluate(embeddings, labels, nrof_folds=10, distance_metric=0, subtract_mean=False,
             thresholds=np.arange(0, 4, 0.01)):
    """utility.py:10-33: embeddings interleaved (even rows = first of each pair)."""
    embeddings = _ffi.host_array(embeddings, np.float32)
    embeddings1 = embeddings[0::2]
    embeddings2 = embeddings[1::2]
    tpr, fpr, accuracy, f1scores = calculate_roc(thresholds, embeddings1, embeddings2, np.asarray(labels),
                                                 nrof_folds=nrof_folds, distance_metric=distance_metric,
                                                 subtract_mean=subtract_mean)
    thresholds = np.arange(0, 4, 0.001)
    far_target = 1e-3
    val, val_std, far = calculate_val(thresholds, embeddings1, embeddings2, np.asarray(labels), far_target,
                                      nrof_folds=nrof_folds, distance_metric=distance_metric,
                                      subtract_mean=subtract_mean)
    return tpr, fpr, accuracy, f1scores, val, val_std, far


def read_pairs(pairs_filename):
    """utility.py:256-262 with a list-of-lists return (np.array of ragged rows fails on numpy >= 1.24)."""
    pairs = []
    with open(pairs_filename, 'r') as f:
        for line in f.readlines()[1:]:
            pairs.append(line.strip().split())
    return pairs


def add_extension(path):
    """utility.py:247-253: the image behind an LFW stem is a .jpg or a .png; anything else is an error."""
    for ext in (".jpg", ".png"):
        if os.path.exists(path + ext):
            return path + ext
    raise RuntimeError('No file "%s" with extension png or jpg.' % path)


def get_paths(lfw_dir, pairs):
    """utility.py:222-244: expand pairs.txt rows into a flat [path0, path1, ...] list and the issame flags.
    A 3-field row is (name, i, j) of one person, a 4-field row (name1, i, name2, j) of two; images are
    `<dir>/<name>/<name>_%04d.{jpg,png}`.  Rows whose images are missing are skipped and counted, as in the
    reference (which, however, raises from add_extension before it can skip: here a missing stem skips too)."""
    path_list, issame_list, skipped = [], [], 0

    def stem(name, idx):
        return os.path.join(lfw_dir, name, "%s_%04d" % (name, int(idx)))

    for pair in pairs:
        if len(pair) == 3:
            stems, same = (stem(pair[0], pair[1]), stem(pair[0], pair[2])), True
        elif len(pair) == 4:
            stems, same = (stem(pair[0], pair[1]), stem(pair[2], pair[3])), False
        else:
            skipped += 1
            continue
        try:
            found = [add_extension(s) for s in stems]
        except RuntimeError:
            skipped += 1
            continue
        path_list += found
        issame_list.append(same)
    if skipped > 0:
        print('Skipped %d image pairs' % skipped)
    return path_list, issame_list
