"""Development aid: main-loop time of the 3xTF32 / bf16 skeleton vs tile width and split count (ArcFace GEMM shapes)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_insight_face_b200 import _ffi

_ffi.init(0)
lib = _ffi.load_library()
ms = C.c_float()
for (M, N, K) in ((512, 10000, 512), (10000, 512, 512)):
    for prec in (0,):
        for bn in (256, 224, 192, 160, 128, 96, 64):
            n_tiles = (N + bn - 1) // bn
            m_blocks = (M + 255) // 256
            for tps in (1, 2, 3, 4):
                splits = (n_tiles + tps - 1) // tps
                if splits * m_blocks > 74 * 3 or tps > n_tiles:
                    continue
                rc = lib.dif_debug_gemm_time(M, N, K, prec, 2 | (bn << 8), splits, 20, C.byref(ms))
                fl = 2.0 * M * N * K
                print(f"M {M} N {N} K {K} prec {prec} bn {bn} tiles {n_tiles} tps {tps} splits {splits} items {splits * m_blocks}: "
                      f"rc {rc} {ms.value * 1e3:8.1f} us  {fl / ms.value / 1e9:7.1f} TFLOP/s", flush=True)
