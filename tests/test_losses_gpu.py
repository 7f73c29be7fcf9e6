"""Batch-hard triplet losses (common/losses.py:33-128) and the row-wise losses through the C ABI against the
numpy oracle: mined indices bit-exact (ties -> lower index), loss / gradient within 1e-4 relative
(BASELINE.json), reference edge cases of SURVEY.md section 8c."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # BASELINE.json: loss, gradients and distances within 1e-4 relative in fp32


def pk_batch(P, K, D, noise, seed=1, normalise=False):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((P, D)).astype(np.float32)
    emb = (np.repeat(cent, K, 0) + noise * rng.standard_normal((P * K, D))).astype(np.float32)
    if normalise:
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    return emb, np.repeat(np.arange(P), K)


def close(got, want, scale=None, rtol=RTOL):
    scale = np.abs(want).max() if scale is None else scale
    assert np.abs(np.asarray(got, dtype=np.float64) - want).max() <= rtol * max(scale, 1e-30), \
        f"max err {np.abs(got - want).max()} vs scale {scale}"


def run_case(cls, oracle_fn, emb, lab, alpha, onehot=True):
    loss = cls(alpha=alpha)
    labels = np.eye(lab.max() + 1, dtype=np.float32)[lab] if onehot else lab
    got, grad, info = loss.loss_and_grad(labels, emb)
    want = oracle_fn(lab, emb, alpha)
    assert np.array_equal(info["pos_idx"], want["pos_idx"]), "mined positives differ"
    assert np.array_equal(info["neg_idx"], want["neg_idx"]), "mined negatives differ"
    close(got, want["loss"], scale=max(np.abs(want["loss"]).max(), np.abs(want["hardest_pos"]).max()))
    # gradient scale floor: one active anchor moves its row by ~ |x| / B (exactly cancelling terms leave rounding noise)
    close(grad, want["grad"], scale=max(np.abs(want["grad"]).max(), np.abs(emb).max() / emb.shape[0]))
    close(info["stats"], want["stats"], scale=np.abs(want["stats"]).max())
    return got, grad, info, want


@pytest.mark.parametrize("P,K,D", [(18, 4, 128), (8, 4, 96), (33, 3, 100), (64, 4, 128), (256, 4, 128), (16, 2, 512)])
@pytest.mark.parametrize("noise", [0.5, 1.5])
def test_batch_hard_cosine(gpu, P, K, D, noise):
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(P, K, D, noise)
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 0.35)


@pytest.mark.parametrize("P,K,D", [(18, 4, 128), (33, 3, 100), (256, 4, 128)])
@pytest.mark.parametrize("alpha,normalise", [(0.35, True), (50.0, False)])
def test_batch_hard_euclidean(gpu, P, K, D, alpha, normalise):
    from deep_insight_face_b200.common.losses import BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(P, K, D, 1.5, normalise=normalise)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, alpha)


def test_batch_hard_large_batch_indices(gpu):
    """B = 4096 (top of the BASELINE sweep): mined indices still bit-exact."""
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(1024, 4, 128, 1.0)
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 0.35, onehot=False)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, 100.0, onehot=False)


def test_edge_cases_single_sample_identity_and_single_identity(gpu):
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(10, 3, 64, 1.0)
    lab = lab.copy()
    lab[-1] = 99  # an identity with K = 1: its only positive is itself (filler 1.0 / 0 may win)
    lab = np.unique(lab, return_inverse=True)[1]
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 0.35)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, 10.0)
    one = np.zeros(30, dtype=np.int64)   # one identity only: hardest negative is the filler (-1 / max(dists))
    _, _, info, _ = run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, one, 0.35)
    assert (info["neg_idx"] == -1).all()
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, one, 0.35)
    _, grad, info, want = run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, one, 500.0)
    assert (info["neg_idx"] == -1).all() and np.abs(want["grad"]).max() > 0  # gradient flows through max(dists)


def test_exact_ties_split_the_gradient(gpu):
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(2)
    emb = rng.standard_normal((12, 32)).astype(np.float32)
    emb[5] = emb[4]
    emb[9] = emb[8]
    emb[2] = 0.0      # a zero row: l2_normalize leaves it at 0
    lab = np.repeat(np.arange(4), 3)
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 5.0)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, 5.0)


def test_keras_protocol_and_auto_alpha(gpu):
    from deep_insight_face_b200.common.losses import (BatchHardTripletLoss, BatchHardTripletLossEuclideanAutoAlpha,
                                                      TripletLossWapper)
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    onehot = np.eye(18, dtype=np.float32)[lab]
    loss = BatchHardTripletLoss(alpha=0.2)
    cfg = loss.get_config()
    assert cfg["alpha"] == 0.2
    again = BatchHardTripletLoss.from_config(cfg)
    per_sample = again.call(onehot, emb)
    assert per_sample.shape == (72,)
    close(per_sample, lo.batch_hard_cosine(lab, emb, 0.2)["loss"])
    close(again(onehot, emb), lo.batch_hard_cosine(lab, emb, 0.2)["loss"].mean())   # Keras AUTO reduction
    assert TripletLossWapper().call(onehot, emb) is None                              # losses.py:17-18
    auto = BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    first = auto.call(onehot, emb)
    want1 = lo.batch_hard_euclidean(lab, emb, 1.0)                                    # uses the PREVIOUS alpha (:112)
    close(first, want1["loss"], scale=np.abs(want1["hardest_pos"]).max())
    assert abs(auto.auto_alpha - want1["stats"][0] * 0.1) <= 1e-4 * want1["stats"][0]  # :113
    second = auto.call(onehot, emb)
    close(second, lo.batch_hard_euclidean(lab, emb, auto.auto_alpha)["loss"], scale=np.abs(want1["hardest_pos"]).max())
    assert "auto_alpha" not in auto.get_config()


def test_torch_autograd_bridge(gpu):
    import torch

    from deep_insight_face_b200.common.losses import BatchHardTripletLoss
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    x = torch.from_numpy(emb).cuda().requires_grad_(True)
    onehot = torch.from_numpy(np.eye(18, dtype=np.float32)[lab]).cuda()
    per_sample = BatchHardTripletLoss().call(onehot, x)
    per_sample.mean().backward()
    want = lo.batch_hard_cosine(lab, emb, 0.35)
    close(x.grad.cpu().numpy(), want["grad"])


def test_explicit_triplet_siamese_and_contrastive(gpu):
    from deep_insight_face_b200.networks.siamese import _accuracy, contrastive_loss, euclidean_distance
    from deep_insight_face_b200.networks.triplet import triplet_loss
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(3)
    y = rng.standard_normal((50, 3 * 128)).astype(np.float32)
    got, grad = triplet_loss(None, y, alpha=0.4, return_grad=True)
    want = lo.triplet_apn(y, 0.4)
    assert np.array_equal(got.view(np.uint32), want["loss"].view(np.uint32))   # canonical arithmetic: bit-exact
    close(grad, want["grad"])
    a, b = y[:, :128], y[:, 128:256]
    d = euclidean_distance([a, b])
    assert d.shape == (50, 1)
    close(d, lo.euclidean_distance(a, b))
    yt = (rng.random(50) > 0.5).astype(np.float32)
    val, dd = contrastive_loss(yt, d / 10, return_grad=True)
    wv, wd = lo.contrastive_loss(yt, d / 10)
    assert abs(val - wv) <= RTOL * abs(wv)
    close(dd, wd)
    assert _accuracy(yt, d[:, 0] / 10, 0.5) == lo.siamese_accuracy(yt, d[:, 0] / 10, 0.5)


@pytest.mark.parametrize("P,K,D", [(18, 4, 128), (33, 3, 100), (128, 4, 256), (16, 2, 512),
                                   (256, 4, 128), (100, 3, 96), (67, 5, 50), (64, 4, 512), (90, 3, 200)])   # B >= 256: matrix path
@pytest.mark.parametrize("noise,alpha", [(0.5, 0.35), (1.5, 0.2)])
def test_batch_all(gpu, P, K, D, noise, alpha):
    from deep_insight_face_b200.common.losses import BatchAllTripletLoss
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(P, K, D, noise)
    onehot = np.eye(P, dtype=np.float32)[lab]
    loss = BatchAllTripletLoss(alpha=alpha)
    got, grad, _ = loss.loss_and_grad(onehot, emb)
    want = lo.batch_all_cosine(lab, emb, alpha)
    close(got, want["loss"])
    close(grad, want["grad"], scale=max(np.abs(want["grad"]).max(), np.abs(emb).max() / emb.shape[0] * 1e-2))
    close(loss.call(onehot, emb), want["loss"])


def test_batch_all_matrix_path_decides_like_the_tile_path(gpu):
    """B >= 256 builds the similarity matrix once (canon_mm.cuh) instead of three tile passes: the same canonical
    products, so the valid-negative counts - hence every loss term that is not a float sum - cannot move; the float
    sums are folded in another order, so losses agree to rounding."""
    from deep_insight_face_b200.common.losses import BatchAllTripletLoss
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(128, 4, 128, 1.0)
    emb[17] = emb[3]          # an exact duplicate
    emb[200] = 0.0            # a zero row: l2_normalize clamps
    got, grad, _ = BatchAllTripletLoss(alpha=0.35).loss_and_grad(lab, emb)
    want = lo.batch_all_cosine(lab, emb, 0.35)
    close(got, want["loss"])
    close(grad, want["grad"], scale=max(np.abs(want["grad"]).max(), np.abs(emb).max() / emb.shape[0] * 1e-2))
    got2, grad2, _ = BatchAllTripletLoss(alpha=0.35).loss_and_grad(lab, emb)
    assert np.array_equal(got, got2) and np.array_equal(grad, grad2)


def test_cuda_losses_match_the_reference_source(gpu):
    """The CUDA path against tests/golden/losses_reference.npz directly: what deep_insight_face/common/losses.py (imported
    and run in the build container on a float64 stand-in for its TensorFlow calls), networks/triplet.py:triplet_loss and
    networks/siamese.py compute, with the gradient of mean(loss) from autograd through the reference's own op sequence."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_losses import CASES, pk_batch as ref_batch

    from deep_insight_face_b200.common.losses import (BatchAllTripletLoss, BatchHardTripletLoss, BatchHardTripletLossEuclidean,
                                                      BatchHardTripletLossEuclideanAutoAlpha)
    from deep_insight_face_b200.networks.siamese import _accuracy, contrastive_loss, euclidean_distance
    from deep_insight_face_b200.networks.triplet import triplet_loss

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_reference.npz"))

    def near(got, want, what, floor):
        got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
        scale = max(np.abs(want).max(), floor)
        assert np.abs(got - want).max() <= RTOL * scale, f"{what}: {np.abs(got - want).max():.3e} vs {scale:.3e}"

    for name, P, K, D, noise, seed, flags in CASES:
        emb, lab = ref_batch(P, K, D, noise, seed, **flags)
        onehot = np.eye(int(lab.max()) + 1, dtype=np.float32)[lab]
        for key, obj in (("bh_cos", BatchHardTripletLoss(alpha=0.35)), ("bh_euc", BatchHardTripletLossEuclidean(alpha=0.3 * D)),
                         ("ball", BatchAllTripletLoss(alpha=0.35))):
            loss, grad, _ = obj.loss_and_grad(onehot, emb)
            near(loss, ref[f"{name}/{key}/loss"], f"{name}/{key} loss", 1.0)
            near(grad, ref[f"{name}/{key}/grad"], f"{name}/{key} grad", 1e-3)
        auto = BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
        for step in range(2):
            loss, grad, _ = auto.loss_and_grad(onehot, emb)
            near(loss, ref[f"{name}/bh_auto{step}/loss"], f"{name}/auto{step} loss", 1.0)
            near(grad, ref[f"{name}/bh_auto{step}/grad"], f"{name}/auto{step} grad", 1e-3)   # (single identity: +g and -g cancel in fp32)
            want_alpha = float(ref[f"{name}/bh_auto{step}/auto_alpha_after"])
            assert abs(float(auto.auto_alpha) - want_alpha) <= RTOL * abs(want_alpha)
    for name in ("apn_a", "apn_b"):
        y = ref[f"{name}/y"]
        loss, grad = triplet_loss(None, y, alpha=0.4, return_grad=True)
        near(loss, ref[f"{name}/loss"], name + " loss", 1.0)
        near(grad / y.shape[0], ref[f"{name}/grad"], name + " grad", 1e-3)
    a, b = ref["siamese/a"], ref["siamese/b"]
    d = euclidean_distance([a, b])
    near(d, ref["siamese/dist"], "euclidean_distance", 1.0)
    y = (np.arange(50) % 2).astype(np.float32)
    val, _ = contrastive_loss(y, d / 10, return_grad=True)
    assert abs(val - float(ref["siamese/contrastive"])) <= RTOL * abs(val)
    assert _accuracy(y, d[:, 0] / 10) == float(ref["siamese/accuracy_default"])


def test_host_step_in_the_staging_block_equals_the_copying_call(gpu):
    """BatchHardHostStep works in the library's page-locked block (dif_batch_hard_host_buffers): no copy-in / copy-out,
    the same bits as loss_and_grad on separate arrays, for both metrics and with a per-sample cotangent."""
    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import BatchHardHostStep, batch_hard

    emb, lab = pk_batch(18, 4, 128, 1.0, seed=5)
    rng = np.random.default_rng(9)
    dl = rng.random(72).astype(np.float32)
    for variant, alpha in ((_ffi.LOSS_BH_COSINE, 0.35), (_ffi.LOSS_BH_EUCLIDEAN, 40.0)):
        for use_dl in (False, True):
            want_loss, want_grad, info = batch_hard(lab, emb, variant, alpha, dloss=dl if use_dl else None)
            step = BatchHardHostStep(72, 128, variant, alpha, use_dloss=use_dl)
            step.emb[:] = emb
            step.labels[:] = lab
            if use_dl:
                step.dloss[:] = dl
            for _ in range(2):
                loss, grad = step()
            assert np.array_equal(loss, want_loss) and np.array_equal(grad, want_grad)
            assert np.array_equal(step.pos_idx, info["pos_idx"]) and np.array_equal(step.neg_idx, info["neg_idx"])
            assert np.array_equal(step.stats, info["stats"])


def test_outgrown_workspaces_can_be_released(gpu):
    """A sweep over batch sizes parks every outgrown workspace block (a captured CUDA graph may still point at it);
    dif_release_retired frees them and the next call allocates afresh."""
    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import BatchAllTripletLoss, BatchHardTripletLoss
    from oracle import losses_oracle as lo

    _ffi.release_retired()
    for P in (130, 160, 200, 260):        # growing B: 520 .. 1040 rows, tensor-core miner and matrix batch-all
        emb, lab = pk_batch(P, 4, 64, 1.0)
        BatchHardTripletLoss().loss_and_grad(lab, emb)
        BatchAllTripletLoss().loss_and_grad(lab, emb)
    assert _ffi.release_retired() > 0
    assert _ffi.release_retired() == 0
    emb, lab = pk_batch(300, 4, 64, 1.0)
    got, grad, _ = BatchHardTripletLoss().loss_and_grad(lab, emb)
    want = lo.batch_hard_cosine(lab, emb, 0.35)
    close(got, want["loss"], scale=1.0)
    close(grad, want["grad"])


def test_batch_hard_step_graph(gpu):
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import BatchHardStep
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    want = lo.batch_hard_cosine(lab, emb, 0.35)
    for graph in (False, True):
        step = BatchHardStep(72, 128, _ffi.LOSS_BH_COSINE, 0.35, "cuda:0", graph=graph)
        step.emb.copy_(torch.from_numpy(emb))
        step.labels.copy_(torch.from_numpy(lab.astype(np.int32)))
        for _ in range(3):
            loss, grad = step()
        torch.cuda.synchronize()
        assert np.array_equal(step.pos_idx.cpu().numpy(), want["pos_idx"])
        close(loss.cpu().numpy(), want["loss"])
        close(grad.cpu().numpy(), want["grad"])


@pytest.fixture
def tensor_path(lib):
    from deep_insight_face_b200 import _ffi

    _ffi.check(lib.dif_batch_hard_set_path(2))
    yield
    _ffi.check(lib.dif_batch_hard_set_path(0))


@pytest.mark.parametrize("P,K,D", [(18, 4, 128), (33, 3, 100), (130, 4, 64), (256, 4, 128), (100, 5, 512)])
@pytest.mark.parametrize("variant", ["cosine", "euclid"])
def test_tensor_core_miner_matches_oracle(gpu, tensor_path, P, K, D, variant):
    """The tcgen05 filter + canonical re-rank path (default for B >= 512), forced on at every size."""
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(P, K, D, 1.0)
    if variant == "cosine":
        run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 0.35, onehot=False)
    else:
        run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, 0.3 * D, onehot=False)


def test_tensor_core_miner_ties_and_single_identity(gpu, tensor_path):
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(2)
    emb = rng.standard_normal((48, 64)).astype(np.float32)
    for a, b in ((5, 4), (9, 8), (10, 8), (11, 8), (12, 8), (13, 8)):   # six copies of row 8: overflows a 4-slot list
        emb[a] = emb[b]
    emb[2] = 0.0
    lab = np.repeat(np.arange(6), 8)
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, lab, 5.0, onehot=False)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, lab, 5.0, onehot=False)
    one = np.zeros(48, dtype=np.int64)
    run_case(BatchHardTripletLoss, lo.batch_hard_cosine, emb, one, 0.35, onehot=False)
    run_case(BatchHardTripletLossEuclidean, lo.batch_hard_euclidean, emb, one, 500.0, onehot=False)


@pytest.mark.parametrize("P,K,D", [(18, 4, 128), (256, 4, 128)])
def test_soft_margin_extension(gpu, P, K, D):
    """log(1 + exp(.)) margin of arXiv 1703.07737 (not in the reference; oracle = this build's restatement)."""
    from deep_insight_face_b200.common.losses import BatchHardTripletLoss, BatchHardTripletLossEuclidean
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(P, K, D, 1.0, normalise=True)
    for cls, fn in ((BatchHardTripletLoss, lo.batch_hard_cosine), (BatchHardTripletLossEuclidean, lo.batch_hard_euclidean)):
        loss = cls(soft=True)
        got, grad, info = loss.loss_and_grad(lab, emb)
        want = fn(lab, emb, soft=True)
        assert np.array_equal(info["pos_idx"], want["pos_idx"]) and np.array_equal(info["neg_idx"], want["neg_idx"])
        close(got, want["loss"])
        close(grad, want["grad"])
        assert (got > 0).all() and cls.from_config(loss.get_config()).soft


@pytest.mark.parametrize("variant", ["cosine", "euclid"])
def test_hub_rows_on_the_tensor_core_path(gpu, lib, variant):
    """A row that is the hardest negative of hundreds of anchors (a sample near the centre of the batch): the gradient
    launch takes its anchors off the shared-memory bitmap 32 per round instead of the 30-entry list.  Forced tensor-core
    path against the oracle and against the CUDA-core path, B = 640 (also with exact duplicates, which switch every row
    to the tie-aware walk)."""
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import batch_hard
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(17)
    B, D = 640, 64
    lab = np.repeat(np.arange(B // 4), 4).astype(np.int32)
    emb = (rng.standard_normal((B // 4, D))[lab] * 3.0 + rng.standard_normal((B, D))).astype(np.float32)
    if variant == "cosine":
        emb += 4.0      # a common offset: every row looks alike in angle, row 5 (the offset direction itself) most of all
        emb[5] = 7.0 + 0.01 * rng.standard_normal(D).astype(np.float32)
    else:
        emb[5] = emb.mean(0) + 0.01 * rng.standard_normal(D).astype(np.float32)   # the centre of the batch
    code = _ffi.LOSS_BH_COSINE if variant == "cosine" else _ffi.LOSS_BH_EUCLIDEAN
    alpha = 0.35 if variant == "cosine" else 500.0
    fn = lo.batch_hard_cosine if variant == "cosine" else lo.batch_hard_euclidean
    for dup in (False, True):
        x = emb.copy()
        if dup:
            x[77] = x[12]
        want = fn(lab, x, alpha)
        hub = np.bincount(want["neg_idx"][want["neg_idx"] >= 0], minlength=B).max()
        assert hub > 64, hub      # the batch really has a hub
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
        out = {}
        try:
            for path in (1, 2):
                _ffi.check(lib.dif_batch_hard_set_path(path))
                loss, grad, info = batch_hard(yd, xd, code, alpha)
                out[path] = (loss.cpu().numpy(), grad.cpu().numpy(), info["neg_idx"].cpu().numpy())
        finally:
            _ffi.check(lib.dif_batch_hard_set_path(0))
        assert np.array_equal(out[2][2], want["neg_idx"]) and np.array_equal(out[1][2], want["neg_idx"])
        close(out[2][0], want["loss"], scale=max(1.0, np.abs(want["loss"]).max()))
        close(out[2][1], want["grad"])
        close(out[2][1], out[1][1], scale=np.abs(out[1][1]).max())


def test_graphed_steps_survive_workspace_growth():
    """A CUDA-graphed step keeps the library workspace addresses in its kernel nodes: a later, larger step makes the
    workspace grow, and the outgrown block must stay valid (retired, not freed) for the earlier graph."""
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import BatchHardStep

    rng = np.random.default_rng(3)
    small = BatchHardStep(640, 128, _ffi.LOSS_BH_COSINE, 0.35, "cuda:0", graph=True)
    emb = rng.standard_normal((640, 128)).astype(np.float32)
    small.emb.copy_(torch.from_numpy(emb))
    small.labels.copy_(torch.from_numpy(np.repeat(np.arange(160), 4).astype(np.int32)))
    small()
    torch.cuda.synchronize()
    before = [t.clone() for t in (small.loss, small.grad)]
    big = BatchHardStep(3072, 128, _ffi.LOSS_BH_COSINE, 0.35, "cuda:0", graph=True)   # grows every workspace
    big()
    junk = torch.full((64 << 20,), 7.0, device="cuda")     # would land in the freed blocks if they had been freed
    small.loss.zero_()
    small.grad.zero_()
    small()
    torch.cuda.synchronize()
    assert torch.equal(small.loss, before[0]) and torch.equal(small.grad, before[1])
    del junk


def test_auto_alpha_is_differentiable_on_device_tensors(gpu):
    """BatchHardTripletLossEuclideanAutoAlpha.call on torch-CUDA embeddings must carry a grad_fn like the
    other three classes, use the PREVIOUS auto_alpha (losses.py:112) and update the state afterwards (:113)."""
    import torch

    from deep_insight_face_b200.common.losses import BatchHardTripletLossEuclideanAutoAlpha
    from oracle import losses_oracle as lo

    emb, lab = pk_batch(18, 4, 128, 1.5)
    auto = BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    x = torch.from_numpy(emb).cuda().requires_grad_(True)
    y = torch.from_numpy(lab).cuda()
    per_sample = auto.call(y, x)
    assert per_sample.requires_grad and per_sample.grad_fn is not None
    per_sample.mean().backward()
    want = lo.batch_hard_euclidean(lab, emb, 1.0)
    close(per_sample.detach().cpu().numpy(), want["loss"], scale=np.abs(want["hardest_pos"]).max())
    close(x.grad.cpu().numpy(), want["grad"])
    assert abs(auto.auto_alpha - want["stats"][0] * 0.1) <= 1e-4 * want["stats"][0]
    x.grad = None
    second = auto.call(y, x)        # now with margin mean(dists) * 0.1
    second.mean().backward()
    want2 = lo.batch_hard_euclidean(lab, emb, float(want["stats"][0]) * 0.1)
    close(second.detach().cpu().numpy(), want2["loss"], scale=np.abs(want2["hardest_pos"]).max())
    close(x.grad.cpu().numpy(), want2["grad"])


class _FakeTensor:
    """The slice of tf.Tensor the bridge touches."""

    def __init__(self, a):
        self.a = np.asarray(a)

    def numpy(self):
        return self.a


class _FakeTF:
    """A stand-in `tensorflow` module with the four entry points _tf_call uses, executed eagerly: Tensor / Variable
    types, float32, custom_gradient (keeps the grad_fn so the test can drive the backward pass) and numpy_function."""

    Tensor = _FakeTensor
    Variable = _FakeTensor
    float32 = np.float32

    def __init__(self):
        self.grad_fns = []
        self.numpy_calls = 0

    def custom_gradient(self, f):
        def wrapped(*args):
            out, grad_fn = f(*args)
            self.grad_fns.append(grad_fn)
            return out

        return wrapped

    def numpy_function(self, fn, inputs, dtype):
        self.numpy_calls += 1
        args = [i.numpy() if isinstance(i, _FakeTensor) else np.asarray(i) for i in inputs]
        return _FakeTensor(np.asarray(fn(*args), dtype=dtype))


def test_tf_bridge_executes_under_a_fake_tensorflow(gpu, monkeypatch):
    """TensorFlow is not installable here, so the tf.custom_gradient + tf.numpy_function bridge (the path Keras `fit`
    takes: reference networks/triplet.py:182,209,211) is driven through a minimal fake module: forward and backward
    must reach the kernels and reproduce the oracle, and AutoAlpha must read its margin when the step runs."""
    from deep_insight_face_b200.common import losses as L
    from oracle import losses_oracle as lo

    fake = _FakeTF()
    monkeypatch.setattr(L, "_tf", fake)
    emb, lab = pk_batch(18, 4, 128, 1.5)
    onehot = np.eye(18, dtype=np.float32)[lab]
    loss = L.BatchHardTripletLoss(alpha=0.2)
    out = loss.call(_FakeTensor(onehot), _FakeTensor(emb))
    assert isinstance(out, _FakeTensor) and fake.numpy_calls == 1
    want = lo.batch_hard_cosine(lab, emb, 0.2)
    close(out.numpy(), want["loss"])
    dl = np.full(72, 1.0 / 72, dtype=np.float32)
    g = fake.grad_fns[-1](_FakeTensor(dl))
    close(g.numpy(), want["grad"])
    auto = L.BatchHardTripletLossEuclideanAutoAlpha(alpha=0.1, init_auto_alpha=1)
    first = auto.call(_FakeTensor(onehot), _FakeTensor(emb))
    want1 = lo.batch_hard_euclidean(lab, emb, 1.0)
    close(first.numpy(), want1["loss"], scale=np.abs(want1["hardest_pos"]).max())
    g1 = fake.grad_fns[-1](_FakeTensor(dl))            # backward of step 1 still uses margin 1.0
    close(g1.numpy(), want1["grad"])
    assert abs(auto.auto_alpha - want1["stats"][0] * 0.1) <= 1e-4 * want1["stats"][0]


@pytest.mark.parametrize("B,D", [(72, 128), (8, 64), (128, 256), (100, 96), (33, 512), (5, 32)])
@pytest.mark.parametrize("variant", ["cosine", "euclid", "euclid_soft"])
def test_one_launch_cluster_step_equals_the_two_launch_path(gpu, lib, B, D, variant):
    """B <= 128 runs as ONE thread-block-cluster launch (csrc/batch_hard.cu:bh_cluster_step_kernel).  It must give the
    same loss bits, mined indices and gradient bits as the CUDA-core miner + merge / gradient launches, and the
    oracle's answer - on a batch with duplicates, a zero row and a singleton identity."""
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.common.losses import batch_hard
    from oracle import losses_oracle as lo

    rng = np.random.default_rng(B + D)
    P = max(1, B // 4)
    lab = (np.arange(B) % P).astype(np.int32)
    lab[-1] = P + 7                                   # an identity with a single sample
    emb = (rng.standard_normal((P + 8, D))[lab % (P + 8)] + 0.7 * rng.standard_normal((B, D))).astype(np.float32)
    if B >= 8:
        emb[3] = emb[1]                               # exact duplicates: tied extremes split the gradient
        emb[6] = 0.0                                  # a zero row
    code = {"cosine": _ffi.LOSS_BH_COSINE, "euclid": _ffi.LOSS_BH_EUCLIDEAN,
            "euclid_soft": _ffi.LOSS_BH_EUCLIDEAN | _ffi.LOSS_SOFT_MARGIN}[variant]
    alpha = 0.35 if variant == "cosine" else 0.3 * D
    x = torch.from_numpy(emb).cuda()
    y = torch.from_numpy(lab).cuda()
    dl = torch.from_numpy(rng.random(B).astype(np.float32)).cuda()
    out = {}
    try:
        for path in (1, 3):
            _ffi.check(lib.dif_batch_hard_set_path(path))
            n0 = _ffi.launch_count()
            loss, grad, info = batch_hard(y, x, code, alpha, dloss=dl)
            torch.cuda.synchronize()
            out[path] = (loss.cpu().numpy(), grad.cpu().numpy(), info["pos_idx"].cpu().numpy(), info["neg_idx"].cpu().numpy(),
                         info["stats"].cpu().numpy(), _ffi.launch_count() - n0)
    finally:
        _ffi.check(lib.dif_batch_hard_set_path(0))
    fits = ((B + 7) // 8 * 8) * D * 4 * (2 if variant == "cosine" else 1) <= 150 * 1024   # the staged batch must fit one SM
    assert out[1][5] >= 2 and out[3][5] == (1 if fits else out[1][5])   # one launch vs miner + merge / gradient
    assert np.array_equal(out[1][0].view(np.uint32), out[3][0].view(np.uint32))
    assert np.array_equal(out[1][2], out[3][2]) and np.array_equal(out[1][3], out[3][3])
    assert np.array_equal(out[1][1].view(np.uint32), out[3][1].view(np.uint32))
    np.testing.assert_allclose(out[3][4], out[1][4], rtol=1e-5)
    fn = lo.batch_hard_cosine if variant == "cosine" else lo.batch_hard_euclidean
    want = fn(lab, emb, alpha, dloss=dl.cpu().numpy(), soft=variant.endswith("soft"))
    assert np.array_equal(out[3][2], want["pos_idx"]) and np.array_equal(out[3][3], want["neg_idx"])
    close(out[3][0], want["loss"], scale=max(1.0, np.abs(want["loss"]).max()))
    close(out[3][1], want["grad"])
