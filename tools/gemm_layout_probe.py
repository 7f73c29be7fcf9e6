"""Development aid: dif_debug_gemm_layout on one shape, prints the error against a float64 matmul."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_insight_face_b200 import _ffi

_ffi.init(0)
lib = _ffi.load_library()
for prec in (0, 3):
    for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        for (M, N, K, bn) in ((64, 64, 320, 128), (512, 512, 1024, 256)):
            torch.manual_seed(1)
            A = torch.randn(M, K, device="cuda")
            B = torch.randn(N, K, device="cuda")
            Ag = A.T.contiguous() if a_mn else A
            Bg = B.T.contiguous() if b_mn else B
            C = torch.full((M, N), float("nan"), device="cuda")
            rc = lib.dif_debug_gemm_layout(_ffi.ptr(Ag), _ffi.ptr(Bg), M, N, K, _ffi.ptr(C), prec, a_mn, b_mn, bn, 1, None)
            torch.cuda.synchronize()
            ref = A.double() @ B.double().T
            scale = (A.double().abs() @ B.double().abs().T).max().item()
            err = (C.double() - ref).abs().max().item() / scale
            nz = float((C != 0).float().mean())
            print(f"prec {prec} a_mn {a_mn} b_mn {b_mn} M {M} N {N} K {K} bn {bn}: rc {rc} {lib.dif_last_error() if rc else ''} rel err {err:.3e} nonzero {nz:.3f}",
                  flush=True)
