"""The CUDA tfa losses against numbers the REFERENCE's own text produced (no oracle in between):

 * TripletSemiHardLoss(distance_metric='squared-L2', margin=1) against tf.contrib's triplet_semihard_loss as the
   reference carries it (common/losses.py:151-308, dangling -2ab line repaired; tests/golden/make_golden_semihard.py);
 * TripletHardLoss(distance_metric='squared-L2', margin=alpha) against the reference's BatchHardTripletLossEuclidean
   (common/losses.py:54-85; tests/golden/make_golden_losses.py): the same rule anchor by anchor when every anchor
   has a negative, so the scalar is the mean of the reference's per-anchor losses and the gradient its gradient.

The goldens are float64; the kernels compute in fp32 on the canonical matrix: 3e-4 of the result's scale."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 3e-4


def _close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1e-3)      # batches whose hinges are all inactive have an all-zero gradient
    err = np.abs(got - want).max()
    assert err <= TOL * scale, f"{what}: max error {err:.3e} vs scale {scale:.3e}"


def test_semihard_matches_the_reference_text(gpu):
    from make_golden_semihard import CASES, case_inputs

    from deep_insight_face_b200.common.tfa_losses import TripletSemiHardLoss

    ref = np.load(os.path.join(HERE, "golden", "semihard_reference.npz"))
    for case in CASES:
        name = case[0]
        emb, lab = case_inputs(*case)
        loss, grad, _ = TripletSemiHardLoss(margin=1.0, distance_metric="squared-L2").loss_and_grad(lab.astype(np.int32), emb)
        want = float(ref[f"{name}/repaired/loss"])
        assert abs(float(loss) - want) <= TOL * max(1.0, abs(want)), (name, float(loss), want)
        _close(grad, ref[f"{name}/repaired/grad"], name + " gradient")


def test_hard_matches_the_reference_euclidean_batch_hard(gpu):
    from make_golden_losses import CASES, pk_batch

    from deep_insight_face_b200.common.tfa_losses import TripletHardLoss

    ref = np.load(os.path.join(HERE, "golden", "losses_reference.npz"))
    for name, P, K, D, noise, seed, flags in CASES:
        if name == "one_identity":      # no anchor has a negative: the two losses fill the empty extreme differently
            continue
        emb, lab = pk_batch(P, K, D, noise, seed, **flags)
        loss, grad, _ = TripletHardLoss(margin=0.3 * D, distance_metric="squared-L2").loss_and_grad(lab.astype(np.int32), emb)
        want = float(ref[f"{name}/bh_euc/loss"].mean())
        assert abs(float(loss) - want) <= TOL * max(1.0, abs(want)), (name, float(loss), want)
        _close(grad, ref[f"{name}/bh_euc/grad"], name + " gradient")
