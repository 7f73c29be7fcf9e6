"""tfa losses with distance_metric='angular' on the GPU (normalise -> squared-L2 step with twice the margin -> half the
loss, gradient back through dif_l2_normalize_bwd; common/tfa_losses.angular_via_squared) against the fp64 shadow
of tfa's literal angular formula (oracle/tfa_oracle.torch_shadow_angular_fp64).  The reference never passes this
metric (networks/triplet.py:196,209,211 use the defaults): an extension, parity unpinned.  fp32 kernels against a
float64 shadow: 3e-4 of the result's scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 3e-4


def _pk(P, K, D, seed):
    rng = np.random.default_rng(seed)
    lab = np.repeat(np.arange(P), K).astype(np.int32)
    x = (rng.normal(size=(P, D))[lab] + 0.7 * rng.normal(size=(P * K, D))).astype(np.float32) * 3.0
    perm = rng.permutation(P * K)
    return lab[perm], x[perm]


@pytest.mark.parametrize("kind", ["hard", "semihard"])
@pytest.mark.parametrize("P,K,D,margin", [(18, 4, 128, 0.7), (9, 3, 64, 0.6), (40, 4, 100, 0.5)])
def test_angular_matches_the_tfa_formula(gpu, kind, P, K, D, margin):
    import torch

    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss, TripletSemiHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(P, K, D, seed=P + K)
    cls = TripletHardLoss if kind == "hard" else TripletSemiHardLoss
    fn = cls(margin=margin, distance_metric="angular")
    want, want_g = orc.torch_shadow_angular_fp64(kind, lab, x, margin)
    scale = max(np.abs(want_g).max(), 1e-6)
    loss, grad, _ = fn.loss_and_grad(lab, x)
    assert abs(loss - want) <= TOL * max(1.0, abs(want)), (loss, want)
    assert np.abs(grad - want_g).max() <= TOL * scale
    assert abs(fn(lab, x) - want) <= TOL * max(1.0, abs(want))
    e = torch.from_numpy(x).cuda().requires_grad_(True)      # the differentiable call on device tensors
    out = fn(torch.from_numpy(lab).cuda(), e)
    (2.0 * out).backward()
    assert abs(float(out.detach()) - want) <= TOL * max(1.0, abs(want))
    assert np.abs(e.grad.cpu().numpy() - 2.0 * want_g).max() <= 2.0 * TOL * scale


def test_soft_angular_matches_the_tfa_formula(gpu):
    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss
    from oracle import tfa_oracle as orc

    lab, x = _pk(18, 4, 128, seed=31)
    want, want_g = orc.torch_shadow_angular_fp64("hard", lab, x, 1.0, soft=True)
    loss, grad, _ = TripletHardLoss(soft=True, distance_metric="angular").loss_and_grad(lab, x)
    assert abs(loss - want) <= TOL * max(1.0, abs(want)), (loss, want)
    assert np.abs(grad - want_g).max() <= TOL * max(np.abs(want_g).max(), 1e-6)

