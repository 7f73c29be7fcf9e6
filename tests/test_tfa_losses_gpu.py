"""tensorflow_addons TripletHardLoss / TripletSemiHardLoss (reference call sites networks/triplet.py:196,209,211)
on the GPU against oracle/tfa_oracle.py: mined columns bit-exact, scalar loss and gradient within 1e-4 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _pk(P, K, D, seed, spread=0.5, scale=1.0):
    rng = np.random.default_rng(seed)
    lab = np.repeat(np.arange(P), K).astype(np.int32)
    cent = rng.normal(size=(P, D)) * scale
    x = (cent[lab] + spread * rng.normal(size=(P * K, D))).astype(np.float32)
    perm = rng.permutation(P * K)
    return lab[perm], x[perm]


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-6)
    return np.abs(a - b).max() <= rtol * scale


@pytest.mark.parametrize("P,K,D,scale", [(18, 4, 128, 0.05), (18, 4, 128, 1.0), (16, 8, 64, 0.1), (33, 3, 100, 0.2),
                                          (64, 4, 512, 0.05), (70, 4, 201, 0.2), (65, 4, 260, 0.1), (90, 3, 30, 0.3)])
@pytest.mark.parametrize("soft", [False, True])
def test_hard_matches_oracle(P, K, D, scale, soft):
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(P, K, D, seed=P * 100 + K, scale=scale)
    want = orc.triplet_hard(lab, x, margin=1.0, soft=soft)
    loss, grad, info = TripletHardLoss(soft=soft).loss_and_grad(lab, x)
    assert np.array_equal(info["pos_idx"], want["pos_idx"])
    assert np.array_equal(info["neg_idx"], want["neg_idx"])
    assert abs(loss - float(want["loss"])) <= RTOL * max(abs(float(want["loss"])), 1e-6)
    l64, g64 = orc.torch_shadow("hard", lab, x, margin=1.0, soft=soft)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6)
    assert _close(grad, g64)


@pytest.mark.parametrize("P,K,D,scale", [(18, 4, 128, 0.05), (18, 4, 128, 1.0), (16, 8, 64, 0.1), (33, 3, 100, 0.2)])
@pytest.mark.parametrize("metric", ["L2", "squared-L2"])
def test_semihard_matches_oracle(P, K, D, scale, metric):
    from deep_insight_face_b200.common.tfa_losses import TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(P, K, D, seed=P * 10 + K, scale=scale)
    sq = metric == "squared-L2"
    want = orc.triplet_semihard(lab, x, margin=1.0, squared=sq)
    loss, grad, _ = TripletSemiHardLoss(distance_metric=metric).loss_and_grad(lab, x)
    assert abs(loss - float(want["loss"])) <= RTOL * max(abs(float(want["loss"])), 1e-6)
    l64, g64 = orc.torch_shadow("semihard", lab, x, margin=1.0, squared=sq)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6)
    assert _close(grad, g64)


def test_edge_cases_single_identity_and_singletons():
    """No negatives at all: hn falls back to the row maximum (and its gradient flows there); identities with one
    sample have no positive: hp = 0, mined index -1; semi-hard with no positive pair is NaN as in tfa."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    rng = np.random.default_rng(5)
    x = rng.normal(size=(24, 64)).astype(np.float32)
    one = np.zeros(24, np.int32)
    want = orc.triplet_hard(one, x)
    loss, grad, info = TripletHardLoss().loss_and_grad(one, x)
    assert np.array_equal(info["neg_idx"], want["neg_idx"]) and (info["neg_idx"] == -1).all()
    assert np.array_equal(info["pos_idx"], want["pos_idx"])
    l64, g64 = orc.torch_shadow("hard", one, x)
    assert abs(loss - l64) <= RTOL * abs(l64) and _close(grad, g64)

    singles = np.arange(24, dtype=np.int32)
    want = orc.triplet_hard(singles, x)
    loss, grad, info = TripletHardLoss().loss_and_grad(singles, x)
    assert (info["pos_idx"] == -1).all() and np.array_equal(info["neg_idx"], want["neg_idx"])
    l64, g64 = orc.torch_shadow("hard", singles, x)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
    loss, _, _ = TripletSemiHardLoss().loss_and_grad(singles, x)
    assert np.isnan(loss) and np.isnan(orc.triplet_semihard(singles, x)["loss"])


def test_duplicate_rows_and_ties():
    """Exact duplicates: distance exactly 0 (error mask), tied extremes resolve to the lower column and split
    the gradient evenly, exactly as reduce_max / reduce_min do."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(8, 4, 64, seed=3, scale=0.1)
    x[5] = x[9]
    x[17] = x[2]
    x[30] = x[2]
    for soft in (False, True):
        want = orc.triplet_hard(lab, x, soft=soft)
        loss, grad, info = TripletHardLoss(soft=soft).loss_and_grad(lab, x)
        assert np.array_equal(info["pos_idx"], want["pos_idx"]) and np.array_equal(info["neg_idx"], want["neg_idx"])
        l64, g64 = orc.torch_shadow("hard", lab, x, soft=soft)
        assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
    loss, grad, _ = TripletSemiHardLoss().loss_and_grad(lab, x)
    l64, g64 = orc.torch_shadow("semihard", lab, x)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)


def test_large_batch_forward_and_determinism():
    """B = 2048 (the C4 sweep range): forward against the anchor-by-anchor oracle, gradient finite and
    run-to-run identical."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(512, 4, 128, seed=11, scale=0.05)
    want = orc.triplet_hard(lab, x)
    loss, grad, info = TripletHardLoss().loss_and_grad(lab, x)
    assert np.array_equal(info["pos_idx"], want["pos_idx"]) and np.array_equal(info["neg_idx"], want["neg_idx"])
    assert abs(loss - float(want["loss"])) <= RTOL * float(want["loss"])
    loss2, grad2, _ = TripletHardLoss().loss_and_grad(lab, x)
    assert loss == loss2 and np.array_equal(grad, grad2) and np.isfinite(grad).all()
    ws = orc.triplet_semihard(lab, x)
    ls, gs, _ = TripletSemiHardLoss().loss_and_grad(lab, x)
    assert abs(ls - float(ws["loss"])) <= RTOL * float(ws["loss"])
    ls2, gs2, _ = TripletSemiHardLoss().loss_and_grad(lab, x)
    assert ls == ls2 and np.array_equal(gs, gs2) and np.isfinite(gs).all()


def test_rows_past_the_list_cap_fall_back_to_dense_coefficients():
    """The backward pass carries each anchor's coefficients as a sorted list of at most 32 (column, weight) pairs;
    an anchor with more non-zeros (large identities in the semi-hard loss: one entry per positive plus its selected
    negatives) is written densely instead.  Mixed batch: two identities of 40, ten of 4."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    rng = np.random.default_rng(21)
    lab = np.concatenate([np.repeat([0, 1], 40), np.repeat(np.arange(2, 12), 4)]).astype(np.int32)
    cent = rng.normal(size=(12, 64)) * 0.3
    x = (cent[lab] + 0.5 * rng.normal(size=(lab.size, 64))).astype(np.float32)
    perm = rng.permutation(lab.size)
    lab, x = lab[perm], x[perm]
    for metric in ("L2", "squared-L2"):
        sq = metric == "squared-L2"
        loss, grad, _ = TripletSemiHardLoss(distance_metric=metric).loss_and_grad(lab, x)
        l64, g64 = orc.torch_shadow("semihard", lab, x, squared=sq)
        assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
        loss2, grad2, _ = TripletSemiHardLoss(distance_metric=metric).loss_and_grad(lab, x)
        assert loss == loss2 and np.array_equal(grad, grad2)
    loss, grad, _ = TripletHardLoss().loss_and_grad(lab, x)
    l64, g64 = orc.torch_shadow("hard", lab, x)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)


def test_gradients_on_the_large_batch_path():
    """B >= 256 takes the one-thread-per-entry pairwise kernel and the list gather: gradients against the fp32-faithful
    autograd shadow (hard at B = 2048; semi-hard at B = 288, where the shadow's O(B^3) tensors still fit)."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(512, 4, 128, seed=12, scale=0.05)
    x[100] = x[7]   # an exact duplicate: zero distance, tied extremes
    for soft in (False, True):
        loss, grad, _ = TripletHardLoss(soft=soft).loss_and_grad(lab, x)
        l64, g64 = orc.torch_shadow("hard", lab, x, soft=soft)
        assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
    lab, x = _pk(72, 4, 96, seed=13, scale=0.1)
    for metric in ("L2", "squared-L2"):
        loss, grad, _ = TripletSemiHardLoss(distance_metric=metric).loss_and_grad(lab, x)
        l64, g64 = orc.torch_shadow("semihard", lab, x, squared=metric == "squared-L2")
        assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)


def test_torch_autograd_and_config_roundtrip():
    import torch

    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(18, 4, 128, seed=2, scale=0.05)
    for cls, kind in ((TripletHardLoss, "hard"), (TripletSemiHardLoss, "semihard")):
        fn = cls.from_config(cls(margin=0.7).get_config())
        assert fn.margin == 0.7 and fn.distance_metric == "L2"
        e = torch.from_numpy(x).cuda().requires_grad_(True)
        out = fn(torch.from_numpy(lab).cuda(), e)
        (3.0 * out).backward()
        l64, g64 = orc.torch_shadow(kind, lab, x, margin=0.7)
        assert abs(float(out.detach()) - l64) <= RTOL * max(abs(l64), 1e-6)
        assert _close(e.grad.cpu().numpy(), 3.0 * g64)
    with pytest.raises(NotImplementedError):
        TripletHardLoss(distance_metric="manhattan")


def test_preallocated_and_graphed_step_equals_the_call():
    """TfaTripletStep (fixed buffers, optionally one CUDA graph launch per step) gives the bits of tfa_triplet, also after
    the library workspace has grown for a larger batch in between (outgrown blocks stay alive for captured graphs)."""
    import torch

    from deep_insight_face_b200.common.tfa_losses import TFA_HARD, TFA_SEMIHARD, TfaTripletStep, tfa_triplet

    lab, x = _pk(18, 4, 128, seed=31, scale=0.1)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lab.astype(np.int32)).cuda()
    for kind in (TFA_HARD, TFA_SEMIHARD):
        want_loss, want_grad, want_info = tfa_triplet(ld, xd, kind, 1.0)
        for graph in (False, True):
            step = TfaTripletStep(72, 128, kind, 1.0, "cuda:0", graph=graph)
            step.emb.copy_(xd)
            step.labels.copy_(ld)
            big_lab, big_x = _pk(128, 4, 128, seed=32)          # grows the workspace (B = 512) between capture and replay
            tfa_triplet(torch.from_numpy(big_lab.astype(np.int32)).cuda(), torch.from_numpy(big_x).cuda(), kind, 1.0)
            for _ in range(3):
                loss, grad = step()
            torch.cuda.synchronize()
            assert torch.equal(loss[0], want_loss) and torch.equal(grad, want_grad)
            if kind == TFA_HARD:
                assert torch.equal(step.pos_idx, want_info["pos_idx"]) and torch.equal(step.neg_idx, want_info["neg_idx"])


def test_hub_row_in_the_list_gather():
    """A sample at the centre of the batch is the closest negative of almost every anchor: its row of the gradient
    gather walks its bitmap 32 anchors per round (the per-row list holds 32 entries).  Hard and semi-hard, B = 320."""
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    rng = np.random.default_rng(23)
    lab = np.repeat(np.arange(80), 4).astype(np.int32)
    x = (3.0 * rng.standard_normal((80, 48))[lab] + rng.standard_normal((320, 48))).astype(np.float32)
    x[9] = x.mean(0) + 0.01 * rng.standard_normal(48).astype(np.float32)
    want = orc.triplet_hard(lab, x)
    assert np.bincount(want["neg_idx"][want["neg_idx"] >= 0], minlength=320).max() > 64
    loss, grad, info = TripletHardLoss().loss_and_grad(lab, x)
    assert np.array_equal(info["neg_idx"], want["neg_idx"]) and np.array_equal(info["pos_idx"], want["pos_idx"])
    l64, g64 = orc.torch_shadow("hard", lab, x)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
    loss, grad, _ = TripletSemiHardLoss().loss_and_grad(lab, x)
    l64, g64 = orc.torch_shadow("semihard", lab, x)
    assert abs(loss - l64) <= RTOL * max(abs(l64), 1e-6) and _close(grad, g64)
