"""Main-loop timing sweep of the NT-GEMM skeleton (checksum epilogue)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deep_insight_face_b200 import _ffi
_ffi.init(0)
lib = _ffi.load_library()
M, N, K = 4096, 1_000_000, 512
if len(sys.argv) > 3:
    M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
flops = 2.0 * M * N * K
for prec, name in ((1, "bf16"), (2, "tf32x1"), (0, "tf32x3")):
    for ctas in (1, 2, 18):
        if ctas == 18 and prec == 0:
            continue
        for splits in (37, 5):
            ms = C.c_float()
            rc = lib.dif_debug_gemm_time(M, N, K, prec, ctas, splits, 5, C.byref(ms))
            if rc:
                print(name, ctas, splits, "error:", _ffi.last_error())
                continue
            print(json.dumps({"prec": name, "ctas": ctas & 15, "ares": ctas >> 4, "splits": splits,
                              "ms": round(ms.value, 3), "tflops": round(flops / ms.value / 1e9, 1)}), flush=True)
