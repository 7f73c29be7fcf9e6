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


@pytest.mark.parametrize("prec", [0, 3])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(64, 64, 304, 128, 1), (512, 512, 1000, 256, 2), (300, 520, 96, 128, 5), (1000, 192, 512, 128, 1),
                                   (257, 1000, 128, 256, 3)])
def test_gemm_operand_majors_and_tile_width(lib, gpu, prec, a_mn, b_mn, shape):
    """MN-major operands (the matrix is handed over as [K][rows]: no transposed copy) and runtime tile widths, in both
    fp32-class two-plane modes, against a float64 matmul (ArcFace backward reads dcos and the weight planes this way)."""
    import torch

    from deep_insight_face_b200 import _ffi

    M, N, K, bn, splits = shape
    al = 8 if prec == 3 else 4          # row pitches of the operand planes are 16-byte multiples
    M, N, K = (M + al - 1) // al * al, (N + al - 1) // al * al, (K + al - 1) // al * al
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=gpu)
    B = torch.randn(N, K, device=gpu)
    Ag = A.T.contiguous() if a_mn else A
    Bg = B.T.contiguous() if b_mn else B
    C = torch.full((M, N), float("nan"), device=gpu)
    _ffi.check(lib.dif_debug_gemm_layout(_ffi.ptr(Ag), _ffi.ptr(Bg), M, N, K, _ffi.ptr(C), prec, a_mn, b_mn, bn, splits, None))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T).max().item()
    rel = (C.double() - ref).abs().max().item() / scale
    assert rel == rel and rel < (2e-6 if prec == 0 else 2e-5), f"rel err {rel}"


@pytest.mark.parametrize("bn", [32, 96, 160, 224])
def test_gemm_narrow_tiles(lib, gpu, bn):
    """Tile widths that are odd multiples of 32 (the epilogue's chunk loop has a tail) with K-major operands."""
    import torch

    from deep_insight_face_b200 import _ffi

    M, N, K = 384, 1000, 160
    torch.manual_seed(bn)
    A = torch.randn(M, K, device=gpu)
    B = torch.randn(N, K, device=gpu)
    C = torch.full((M, N), float("nan"), device=gpu)
    _ffi.check(lib.dif_debug_gemm_layout(_ffi.ptr(A), _ffi.ptr(B), M, N, K, _ffi.ptr(C), 0, 0, 0, bn, 3, None))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T).max().item()
    assert (C.double() - ref).abs().max().item() / scale < 2e-6
