"""The tcgen05 / TMA NT-GEMM skeleton behind every dense kernel, against a float64 matmul."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {0: 2e-6, 1: 8e-3, 2: 2e-3}  # 3xTF32 (fp32-exact), bf16, 1xTF32: max |err| / max(|A||B|^T)


@pytest.mark.parametrize("prec,ctas", [(0, 1), (0, 2), (1, 1), (1, 2), (1, 18), (2, 1), (2, 2), (2, 18)])
@pytest.mark.parametrize("shape", [(128, 256, 64, 1), (300, 1000, 96, 1), (513, 4099, 256, 3), (257, 70000, 128, 37)])
def test_nt_gemm(lib, gpu, prec, ctas, shape):
    import torch

    from deep_insight_face_b200 import _ffi

    M, N, K, splits = shape
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device=gpu)
    B = torch.randn(N, K, device=gpu)
    C = torch.full((M, N), float("nan"), device=gpu)
    _ffi.check(lib.dif_debug_nt_gemm(_ffi.ptr(A), _ffi.ptr(B), M, N, K, _ffi.ptr(C), prec, ctas, splits, None))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T).max().item()
    rel = (C.double() - ref).abs().max().item() / scale
    assert rel == rel and rel < TOL[prec], f"rel err {rel}"
    if prec == 0:  # the hi*hi-only product would sit near 3e-5: make sure all three passes ran
        assert rel < 2e-6
