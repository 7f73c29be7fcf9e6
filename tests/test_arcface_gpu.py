"""ArcFace margin logits + softmax-CE (BASELINE config C2) through the C ABI against the fp64 oracle.
PARITY UNPINNED by the reference (it has no ArcFace); the oracle is this build's specification."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def data(B, C, D, seed=2, wscale=0.01):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, D)).astype(np.float32)
    W = (wscale * rng.standard_normal((C, D))).astype(np.float32)
    y = rng.integers(0, C, size=B)
    # make some samples close to their class centre so both branches of phi are exercised
    X[: B // 4] = (W[y[: B // 4]] / wscale + 0.3 * X[: B // 4]).astype(np.float32)
    X[B // 4: B // 4 + 4] = -W[y[B // 4: B // 4 + 4]] / wscale   # theta + m > pi: easy-margin fallback
    return X, W, y


@pytest.mark.parametrize("precision", ["tf32x3", "bf16x3"])
@pytest.mark.parametrize("B,C,D", [(64, 300, 64), (130, 1000, 128), (512, 10000, 512), (72, 18, 128), (33, 1027, 96)])
def test_arcface_forward_backward(gpu, B, C, D, precision):
    """Both fp32-class operand splits (TF32 hi/lo planes, bf16 b0/b1 planes) against the fp64 oracle at 1e-4."""
    from deep_insight_face_b200.arcface import arcface_loss
    from oracle import losses_oracle as lo

    X, W, y = data(B, C, D)
    loss, dX, dW = arcface_loss(X, W, y, 64.0, 0.5, precision=precision)

    want = lo.arcface(X, W, y, 64.0, 0.5)
    assert np.abs(loss - want["loss"]).max() <= RTOL * np.abs(want["loss"]).max()
    assert np.abs(dX - want["dX"]).max() <= RTOL * np.abs(want["dX"]).max()
    assert np.abs(dW - want["dW"]).max() <= RTOL * np.abs(want["dW"]).max()
    fwd_only = arcface_loss(X, W, y, 64.0, 0.5, want_grad=False, precision=precision)
    assert np.array_equal(fwd_only, loss)


def test_arcface_device_entry_and_dloss(gpu):
    import torch

    from deep_insight_face_b200.arcface import ArcFaceLoss, arcface_loss
    from oracle import losses_oracle as lo

    X, W, y = data(96, 500, 128, seed=5)
    xd, wd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(y).cuda()
    loss, dX, dW = arcface_loss(xd, wd, yd, 30.0, 0.35)
    want = lo.arcface(X, W, y, 30.0, 0.35)
    assert np.abs(loss.cpu().numpy() - want["loss"]).max() <= RTOL * np.abs(want["loss"]).max()
    assert np.abs(dW.cpu().numpy() - want["dW"]).max() <= RTOL * np.abs(want["dW"]).max()
    dl = torch.full((96,), 2.0 / 96, device="cuda")
    _, dX2, _ = arcface_loss(xd, wd, yd, 30.0, 0.35, dloss=dl)
    assert torch.allclose(dX2, 2 * dX, rtol=1e-5, atol=1e-9)
    head = ArcFaceLoss(W, 30.0, 0.35)
    onehot = np.eye(500, dtype=np.float32)[y]
    assert abs(head(onehot, X) - want["loss"].mean()) <= RTOL * want["loss"].mean()


def test_arcface_step_graph(gpu):
    import torch

    from deep_insight_face_b200.arcface import ArcFaceStep
    from oracle import losses_oracle as lo

    X, W, y = data(64, 300, 64)
    want = lo.arcface(X, W, y, 64.0, 0.5)
    for graph in (False, True):
        step = ArcFaceStep(64, 300, 64, 64.0, 0.5, "cuda:0", graph=graph)
        step.X.copy_(torch.from_numpy(X))
        step.W.copy_(torch.from_numpy(W))
        step.y.copy_(torch.from_numpy(y.astype(np.int32)))
        for _ in range(3):
            loss, dX, dW = step()
        torch.cuda.synchronize()
        assert np.abs(loss.cpu().numpy() - want["loss"]).max() <= RTOL * np.abs(want["loss"]).max()
        assert np.abs(dW.cpu().numpy() - want["dW"]).max() <= RTOL * np.abs(want["dW"]).max()
