"""One configuration of the NT-GEMM main-loop timing diagnostic: gemm_one.py M N K prec ctas splits iters"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deep_insight_face_b200 import _ffi
_ffi.init(0)
lib = _ffi.load_library()
M, N, K, prec, ctas, splits, iters = [int(x) for x in sys.argv[1:8]]
ms = C.c_float()
_ffi.check(lib.dif_debug_gemm_time(M, N, K, prec, ctas, splits, iters, C.byref(ms)))
print("ms", ms.value, "tflops", 2.0 * M * N * K / ms.value / 1e9)
