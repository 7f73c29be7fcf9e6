"""Pins that need no GPU: host-side code against fixtures produced by the REFERENCE's own functions
(tests/golden/host_reference.json, made by tests/golden/make_golden_host.py in the build container), the mirrored
default arguments, and the ArcFace oracle against an independently written torch fp64 autograd formulation."""
import inspect
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_ref():
    with open(os.path.join(HERE, "golden", "host_reference.json")) as f:
        return json.load(f)


def test_sample_people_reproduces_the_reference_batches(host_ref):
    """generator.py:15-41 executed from the reference's source under np.random.seed(s) vs datagen.sample_people."""
    from make_golden_host import dataset

    from deep_insight_face_b200 import datagen

    for case in host_ref["sample_people"]:
        ds = dataset(case["dataset_seed"], case["n_people"])
        np.random.seed(case["seed"])
        paths, counts = datagen.sample_people(ds, case["P"], case["K"])
        assert paths == case["image_paths"]
        assert [int(c) for c in counts] == case["num_per_class"]
    # a private generator leaves numpy's global stream alone
    np.random.seed(7)
    before = np.random.get_state()[1].copy()
    datagen.sample_people(dataset(100, 40), 18, 4, rng=np.random.default_rng(0))
    assert np.array_equal(before, np.random.get_state()[1])


def test_pairs_txt_is_byte_identical_to_the_reference_writer(host_ref, tmp_path):
    from deep_insight_face_b200 import datagen
    from deep_insight_face_b200.evaluation import utility as U

    ref = host_ref["pairs_txt"]
    fn = tmp_path / "pairs.txt"
    datagen.write_pairs_to_file(str(fn), [[tuple(m) for m in f] for f in ref["matches"]],
                                [[tuple(m) for m in f] for f in ref["mismatches"]], ref["num_folds"], ref["n"])
    assert fn.read_bytes() == ref["text"].encode("utf-8")
    pairs = U.read_pairs(str(fn))                      # evaluation/utility.py:256-262 (header skipped)
    assert len(pairs) == 12 and list(pairs[0]) == ["p0_0", "1", "2"] and list(pairs[3]) == ["p0_0", "1", "q0_0", "7"]


def test_mirrored_defaults_equal_the_reference(host_ref):
    from deep_insight_face_b200 import api, predictions
    from deep_insight_face_b200.common import losses
    from deep_insight_face_b200.evaluation import utility
    from deep_insight_face_b200.networks import siamese, triplet

    d = host_ref["defaults"]

    def dflt(fn, name):
        return inspect.signature(fn).parameters[name].default

    assert dflt(losses.TripletLossWapper.__init__, "alpha") == d["TripletLossWapper.__init__"]["alpha"] == 0.35
    auto = d["BatchHardTripletLossEuclideanAutoAlpha.__init__"]
    assert dflt(losses.BatchHardTripletLossEuclideanAutoAlpha.__init__, "alpha") == auto["alpha"]
    assert dflt(losses.BatchHardTripletLossEuclideanAutoAlpha.__init__, "init_auto_alpha") == auto["init_auto_alpha"]
    assert dflt(triplet.triplet_loss, "alpha") == d["triplet_loss"]["alpha"]
    assert dflt(siamese._accuracy, "threshold") == d["_accuracy"]["threshold"] == 0.4     # was 0.5
    assert dflt(predictions.TripletPrediction.verify, "threshold") == d["TripletPrediction.verify"]["threshold"]
    assert dflt(predictions.SiamesePrediction.verify, "threshold") == d["SiamesePrediction.verify"]["threshold"]
    assert dflt(api.compare_faces, "tolerance") == d["compare_faces"]["tolerance"]
    assert dflt(utility.evaluate, "nrof_folds") == d["evaluate"]["nrof_folds"]
    assert dflt(utility.evaluate, "distance_metric") == d["evaluate"]["distance_metric"]
    assert dflt(utility.evaluate, "subtract_mean") == d["evaluate"]["subtract_mean"]
    assert dflt(utility.distance, "distance_metric") == d["distance"]["distance_metric"]


def test_siamese_accuracy_default_threshold():
    from deep_insight_face_b200.networks.siamese import _accuracy
    from oracle import losses_oracle as lo

    d = np.array([0.1, 0.39, 0.41, 0.45, 0.7], dtype=np.float32)
    y = np.array([1, 1, 1, 0, 0], dtype=np.float32)
    assert _accuracy(y, d) == lo.siamese_accuracy(y, d) == 0.8          # threshold 0.4: 0.41 is a miss, 0.45 is right
    assert _accuracy(y, d, 0.5) == 0.8 and _accuracy(y, d, 0.42) == 1.0


def _arcface_torch(X, W, y, s, m):
    """Independent formulation: cos(theta + m) through acos, autograd for every derivative, fp64."""
    import torch

    x = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    w = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    lab = torch.tensor(y, dtype=torch.int64)
    xn = x / torch.sqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12))
    wn = w / torch.sqrt(torch.clamp((w * w).sum(1, keepdim=True), min=1e-12))
    cos = torch.clamp(xn @ wn.T, -1.0, 1.0)
    cy = cos[torch.arange(len(y)), lab]
    theta = torch.acos(torch.clamp(cy, -1 + 1e-15, 1 - 1e-15))
    margin = torch.where(theta + m < math.pi, torch.cos(theta + m), cy - m * math.sin(math.pi - m))
    onehot = torch.nn.functional.one_hot(lab, W.shape[0]).bool()
    logits = s * torch.where(onehot, margin[:, None], cos)
    loss = -(torch.log_softmax(logits, dim=1)[torch.arange(len(y)), lab])
    loss.mean().backward()
    return loss.detach().numpy(), x.grad.numpy(), w.grad.numpy()


@pytest.mark.parametrize("s,m", [(64.0, 0.5), (30.0, 0.35), (16.0, 1.2)])
def test_arcface_oracle_against_torch_autograd(s, m):
    """the ArcFace oracle had no independent cross-check.  Includes rows past the easy-margin
    switch (theta + m > pi: the sample points away from its class centre) and a non-target cosine that rounds above 1
    (clip edge: tf.clip_by_value / torch.clamp pass no gradient outside the bounds)."""
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(11)
    B, C, D = 48, 37, 24
    X = rng.standard_normal((B, D))
    W = rng.standard_normal((C, D))
    y = rng.integers(0, C, size=B)
    for i in range(0, 12):           # easy-margin branch: x close to -w_y
        X[i] = -W[y[i]] * (0.5 + i) + 0.05 * rng.standard_normal(D)
    for i in range(12, 18):          # near the switch from both sides
        ang = math.pi - m + (i - 14.5) * 0.02
        u = W[y[i]] / np.linalg.norm(W[y[i]])
        v = rng.standard_normal(D)
        v -= u * (u @ v)
        v /= np.linalg.norm(v)
        X[i] = 3.0 * (math.cos(ang) * u + math.sin(ang) * v)
    X[20] = 2.5 * W[(y[20] + 1) % C]  # exactly parallel to a NON-target centre: raw cosine may round to > 1
    want = lo.arcface(X, W, y, s, m)
    loss, dX, dW = _arcface_torch(X, W, y, s, m)
    easy = want["cos_target"] <= math.cos(math.pi - m)
    assert easy[:12].all() and easy[12:18].any() and not easy[12:18].all()
    np.testing.assert_allclose(want["loss"], loss, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(want["dX"], dX, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(want["dW"], dW, rtol=1e-7, atol=1e-10)


def test_arcface_oracle_without_margin_is_cosine_softmax_cross_entropy():
    """m = 0 removes the margin: the ArcFace loss must then be the library cross-entropy (torch's fused
    F.cross_entropy) of the scaled cosine logits, and its gradients autograd's through F.normalize - a second anchor
    that shares no line with the oracle or with the acos formulation above."""
    import torch
    import torch.nn.functional as F

    from oracle import losses_oracle as lo

    rng = np.random.default_rng(12)
    B, C, D, s = 40, 29, 16, 30.0
    X, W, y = rng.standard_normal((B, D)), rng.standard_normal((C, D)), rng.integers(0, C, size=B)
    want = lo.arcface(X, W, y, s, 0.0)
    x = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    w = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    loss = F.cross_entropy(s * (F.normalize(x, dim=1) @ F.normalize(w, dim=1).T), torch.tensor(y), reduction="none")
    loss.mean().backward()
    np.testing.assert_allclose(want["loss"], loss.detach().numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(want["dX"], x.grad.numpy(), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(want["dW"], w.grad.numpy(), rtol=1e-7, atol=1e-10)


# ------------------------------------------------------------------ losses: the reference's own source under an op stand-in
@pytest.fixture(scope="module")
def losses_ref():
    return np.load(os.path.join(HERE, "golden", "losses_reference.npz"))


def _close(got, want, rtol, what, floor=1e-4):
    """|got - want| <= rtol * max(max |want|, floor).  Losses are O(1) quantities (cosines, margins): floor 1, so that a
    loss that is zero up to fp32 rounding (all-singleton batches) is not held to 1e-5 of nothing; gradients get a
    floor a tenth of their usual size (single-identity and all-singleton batches have analytically zero gradients)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), floor)
    err = np.abs(got - want).max()
    assert err <= rtol * scale, f"{what}: max error {err:.3e} vs scale {scale:.3e}"


def test_losses_oracle_matches_the_reference_source(losses_ref):
    """tests/golden/losses_reference.npz holds what deep_insight_face/common/losses.py itself computes (the module is
    imported and its classes called, on a float64 torch stand-in for the TensorFlow entry points it uses; gradients
    of mean(loss) by autograd through the reference's own op sequence).  The oracle - the checker of every GPU test -
    must agree: losses to fp32 rounding, gradients to 1e-5 of their scale, on PK batches incl. exact duplicates (tied
    extremes split the cotangent), a zero row (l2_normalize clamp), a single identity and all-singleton batches."""
    from make_golden_losses import CASES, pk_batch

    from oracle import losses_oracle as lo

    for name, P, K, D, noise, seed, flags in CASES:
        emb, lab = pk_batch(P, K, D, noise, seed, **flags)
        for key, got in (("bh_cos", lo.batch_hard_cosine(lab, emb, 0.35)),
                         ("bh_euc", lo.batch_hard_euclidean(lab, emb, 0.3 * D)),
                         ("ball", lo.batch_all_cosine(lab, emb, 0.35))):
            _close(got["loss"], losses_ref[f"{name}/{key}/loss"], 2e-5, f"{name}/{key} loss", floor=1.0)
            _close(got["grad"], losses_ref[f"{name}/{key}/grad"], 2e-5, f"{name}/{key} grad")
        # AutoAlpha (losses.py:88-128): step 0 runs with the initial margin 1, then margin = mean(dists) * alpha
        margin = 1.0
        for step in range(2):
            got = lo.batch_hard_euclidean(lab, emb, margin)
            _close(got["loss"], losses_ref[f"{name}/bh_auto{step}/loss"], 2e-5, f"{name}/auto{step} loss", floor=1.0)
            _close(got["grad"], losses_ref[f"{name}/bh_auto{step}/grad"], 2e-5, f"{name}/auto{step} grad")
            margin = float(got["stats"][0]) * 0.1
            assert abs(margin - float(losses_ref[f"{name}/bh_auto{step}/auto_alpha_after"])) <= 1e-5 * abs(margin)


def test_apn_and_siamese_oracles_match_the_reference_source(losses_ref):
    """networks/triplet.py:16-46 and networks/siamese.py:22-45, source cut out of the reference and executed."""
    from oracle import losses_oracle as lo

    for name in ("apn_a", "apn_b"):
        got = lo.triplet_apn(losses_ref[f"{name}/y"], 0.4)
        _close(got["loss"], losses_ref[f"{name}/loss"], 2e-5, name + " loss", floor=1.0)
        _close(got["grad"] / got["loss"].shape[0], losses_ref[f"{name}/grad"], 2e-5, name + " grad")   # mean reduction
    a, b = losses_ref["siamese/a"], losses_ref["siamese/b"]
    d = lo.euclidean_distance(a, b)
    _close(d, losses_ref["siamese/dist"], 2e-5, "euclidean_distance")
    assert d[7, 0] == np.float32(np.sqrt(np.float32(1e-7)))            # the K.epsilon() clamp on a zero distance
    y = (np.arange(50) % 2).astype(np.float64)
    cl, dcl = lo.contrastive_loss(y, d / 10.0)
    assert abs(cl - float(losses_ref["siamese/contrastive"])) <= 2e-5 * abs(cl)
    assert lo.siamese_accuracy(y, d / 10.0) == float(losses_ref["siamese/accuracy_default"])


def test_scalar_helpers_match_the_reference_module(losses_ref):
    """networks/utils.py of the reference (imported as is by the golden generator) vs the drop-in's own-words versions."""
    from deep_insight_face_b200.networks import utils as U

    e = losses_ref["api/enc"]
    got = np.array([U.distance(e[i], e[j]) for i in range(6) for j in range(6)], dtype=np.float64)
    assert np.allclose(got, losses_ref["utils/distance"], rtol=1e-6, atol=1e-7)
    d = losses_ref["utils/dgrid"]
    assert np.array_equal(U.distance_to_proba(d), losses_ref["utils/distance_to_proba"])
    assert np.array_equal(U.gaussian_kernel_dist_to_prob(d), losses_ref["utils/gaussian_kernel_1"])
    assert np.array_equal(U.gaussian_kernel_dist_to_prob(d, 0.5), losses_ref["utils/gaussian_kernel_half"])
    assert abs(U.calc_mean_score(losses_ref["utils/scores"]) - float(losses_ref["utils/calc_mean_score"])) <= 1e-12


# ------------------------------------------------------------------ a8: the tfa oracle against the reference's tf.contrib text
@pytest.fixture(scope="module")
def semihard_ref():
    return np.load(os.path.join(HERE, "golden", "semihard_reference.npz"))


def test_masked_extremes_match_the_reference_source(semihard_ref):
    """deep_insight_face/common/losses.py:211-246 (masked_maximum / masked_minimum, imported and called by
    tests/golden/make_golden_semihard.py) are the functions tensorflow_addons kept as _masked_maximum / _masked_minimum;
    the oracle's restatements compute the same numbers, empty masks and one-signed rows included."""
    from oracle import tfa_oracle as t

    data, mask = semihard_ref["masked/data"], semihard_ref["masked/mask"].astype(np.float64)
    assert np.array_equal(t._masked_maximum(data, mask), semihard_ref["masked/maximum"])
    assert np.array_equal(t._masked_minimum(data, mask), semihard_ref["masked/minimum"])


def test_semihard_oracle_matches_the_reference_text_as_it_is(semihard_ref):
    """triplet_loss_adapted_from_tf (common/losses.py:249-308) executed unmodified.  Its pairwise_distance loses the
    -2ab term (line 183 is a discarded expression), so the matrix it ranks is |a|^2 + |b|^2 off the diagonal; on
    that same matrix the oracle's semi-hard rule (anchor by anchor, fp32 and fp64) and the literal tiled forward the
    gradient oracle differentiates give the reference's loss, and autograd through the tiled forward gives the
    reference's gradient."""
    import torch
    from make_golden_semihard import CASES, case_inputs

    from oracle import tfa_oracle as t

    for case in CASES:
        name = case[0]
        emb, lab = case_inputs(*case)
        B = emb.shape[0]
        want_P = semihard_ref[f"{name}/as_is/pdist_squared"]
        sq32 = (emb * emb).sum(1, dtype=np.float32)
        P32 = ((sq32[:, None] + sq32[None, :]) * (1 - np.eye(B, dtype=np.float32))).astype(np.float32)
        _close(P32, want_P, 2e-6, name + " matrix as is")
        want = float(semihard_ref[f"{name}/as_is/loss"])
        got64 = t.semihard_from_matrix(want_P, lab, 1.0)
        assert abs(float(got64["loss"]) - want) <= 1e-12 * max(1.0, abs(want)), name
        assert abs(float(t.semihard_from_matrix(P32, lab, 1.0)["loss"]) - want) <= 2e-5 * max(1.0, abs(want)), name
        x = torch.tensor(emb.astype(np.float64), requires_grad=True)
        sq = (x * x).sum(1, keepdim=True)
        P = (sq + sq.t()) * (1.0 - torch.eye(B, dtype=torch.float64))
        loss = t._torch_loss("semihard", P, torch.tensor(lab), 1.0, False)
        loss.backward()
        assert abs(float(loss.detach()) - want) <= 1e-12 * max(1.0, abs(want)), name
        _close(x.grad.numpy(), semihard_ref[f"{name}/as_is/grad"], 1e-9, name + " gradient as is")


def test_tfa_oracle_matches_the_repaired_reference_text(semihard_ref):
    """The same module with the dangling `- 2.0 * matmul` line joined to its statement (one text repair, asserted to
    match once by the generator) is tf.contrib's triplet_semihard_loss as published - the function tensorflow_addons
    ported.  The oracle's pairwise_distance (squared and not: clamp, error mask, zero diagonal), its semi-hard loss
    with distance_metric='squared-L2', margin 1, and both gradient oracles agree with it: fp64 shadow to 1e-9, the
    canonical fp32 forward to fp32 rounding."""
    from make_golden_semihard import CASES, case_inputs

    from oracle import tfa_oracle as t

    for case in CASES:
        name = case[0]
        emb, lab = case_inputs(*case)
        _close(t.pairwise_distance(emb, True), semihard_ref[f"{name}/repaired/pdist_squared"], 4e-6, name + " squared matrix")
        # sqrt near zero amplifies the rounding of d^2: duplicates are exactly 0 on both sides, the rest is off zero
        _close(t.pairwise_distance(emb, False), semihard_ref[f"{name}/repaired/pdist"], 4e-6, name + " matrix")
        want, want_g = float(semihard_ref[f"{name}/repaired/loss"]), semihard_ref[f"{name}/repaired/grad"]
        got = float(t.triplet_semihard(lab, emb, 1.0, squared=True)["loss"])
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (name, got, want)
        l64, g64 = t.torch_shadow_fp64("semihard", lab, emb, 1.0, False, True)
        assert abs(l64 - want) <= 1e-12 * max(1.0, abs(want)), name
        _close(g64, want_g, 1e-9, name + " fp64 gradient")
        l32, g32 = t.torch_shadow("semihard", lab, emb, 1.0, False, True)
        assert abs(l32 - want) <= 2e-5 * max(1.0, abs(want)), name
        _close(g32, want_g, 2e-5, name + " fp32-faithful gradient")


# ------------------------------------------------------------------ a15: 1:N ranking against the reference's distance function
def test_gallery_oracle_ranks_like_the_reference_distance():
    """The reference has no 1:N search; a search by ITS distance (evaluation/utility.py:52-66, imported and run on
    every (query, gallery row) pair by tests/golden/make_golden_gallery.py) keeps the k smallest.  The oracle's
    gallery_search returns those rows in that order and the same numbers: squared-L2 scores are the reference's
    metric-0 distances, cosine scores are cos(pi * metric-1 distance)."""
    from synth import GALLERY_CASES, check_ranking, gallery_case

    from oracle import c_oracle as orc

    ref = np.load(os.path.join(HERE, "golden", "gallery_reference.npz"))
    for name, seed, N, Q, D, k in GALLERY_CASES:
        rows, q, pick = gallery_case(seed, N, Q, D)
        chk = np.array([rows.astype(np.float64).sum(), q.astype(np.float64).sum(), float(pick.sum())])
        assert np.array_equal(chk, ref[f"{name}/checksum"]), "the inputs are not the ones the golden was made from"
        for metric in (0, 1):
            s, r = orc.gallery_search(rows, q, k, metric)
            share = check_ranking(ref[f"{name}/metric{metric}/dist"], ref[f"{name}/metric{metric}/rows"], s, r, metric)
            assert share > 0.97, (name, metric, share)
            assert (r[:, 0] == pick).all()


# ------------------------------------------------------------------ f1: pairs.txt expansion and one-hot labels
def test_pair_listings_and_one_hot_labels_match_the_reference(host_ref, tmp_path, capsys):
    """generator.py:43-124 (facematch_image_pairs / triplet_image_pairs / create_pairs, cut out of the reference and
    executed by make_golden_host.py on the image tree rebuilt here): same pairs in the same order, same triplets
    under the same np.random.seed, same class names, same name -> one-hot row mapping."""
    from make_golden_host import PAIR_ROWS, image_tree

    from deep_insight_face_b200 import datagen

    ref = host_ref["pair_listing"]
    root = str(tmp_path / "tree")
    os.makedirs(root)
    pairs_txt = image_tree(root)
    rel = lambda p: os.path.relpath(p, root)

    pairs, names = datagen.facematch_image_pairs(root, PAIR_ROWS)
    assert [[rel(a), rel(b), s] for a, b, s in pairs] == ref["facematch"]["pairs"]
    assert sorted(names) == ref["facematch"]["names"]
    for case in ref["triplet"]:
        np.random.seed(case["seed"])
        trip, names = datagen.triplet_image_pairs(root, PAIR_ROWS)
        assert [[rel(x) for x in t] for t in trip] == case["triplets"], case["seed"]
        assert sorted(names) == case["names"]
    for key, func in (("create_facematch", datagen.facematch_image_pairs), ("create_triplet", datagen.triplet_image_pairs)):
        np.random.seed(5)
        pairs, names, one_hot = datagen.create_pairs(root, func=func, pairs_txt=pairs_txt)
        assert len(pairs) == ref[key]["n_pairs"] and one_hot.dtype == np.float32
        assert {n: one_hot[i].tolist() for i, n in enumerate(names)} == ref[key]["one_hot_of"]
    # a missing image skips the row and says so (the reference aborts with add_extension's RuntimeError instead)
    pairs, _ = datagen.facematch_image_pairs(root, [["ann", "1", "9"], ["ann", "1", "2"]])
    assert len(pairs) == 1 and "Skipped 1 image pairs" in capsys.readouterr().out
    with pytest.raises(AssertionError):
        datagen.create_pairs(root, pairs_txt=pairs_txt)


def test_tfa_hard_oracle_meets_the_reference_euclidean_batch_hard_goldens(losses_ref):
    """a8 hard rule: tfa's TripletHardLoss(distance_metric='squared-L2', margin=alpha) is, anchor by anchor, the
    reference's BatchHardTripletLossEuclidean (common/losses.py:54-85) whenever a negative exists for every anchor
    (the two differ only in what fills an empty extreme).  So the goldens the reference's own class produced
    (losses_reference.npz, make_golden_losses.py) pin the tfa oracle's hard rule directly: scalar loss = mean of the
    reference's per-anchor losses, gradient = the reference's gradient of mean(loss) - fp64 shadow to 1e-12 / 1e-9,
    the canonical fp32 forward to fp32 rounding.  'one_identity' is left out: no anchor has a negative there."""
    from make_golden_losses import CASES, pk_batch

    from oracle import tfa_oracle as t

    for name, P, K, D, noise, seed, flags in CASES:
        if name == "one_identity":
            continue
        emb, lab = pk_batch(P, K, D, noise, seed, **flags)
        margin = 0.3 * D
        want, want_g = float(losses_ref[f"{name}/bh_euc/loss"].mean()), losses_ref[f"{name}/bh_euc/grad"]
        l64, g64 = t.torch_shadow_fp64("hard", lab, emb, margin, False, True)
        assert abs(l64 - want) <= 1e-12 * max(1.0, abs(want)), name
        _close(g64, want_g, 1e-9, name + " fp64 gradient")
        got = float(t.triplet_hard(lab, emb, margin=margin, squared=True)["loss"])
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (name, got, want)
        l32, g32 = t.torch_shadow("hard", lab, emb, margin, False, True)
        assert abs(l32 - want) <= 2e-5 * max(1.0, abs(want)), name
        _close(g32, want_g, 2e-5, name + " fp32-faithful gradient")
