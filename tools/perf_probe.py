"""Timing probe of Gallery.search on a synthetic gallery (development aid; bench.py is the contract).

    python tools/perf_probe.py [N] [Q] [D] [k]
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.gallery import Gallery
    from oracle import c_oracle as orc

    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    modes = sys.argv[5].split(",") if len(sys.argv) > 5 else ["tf32x3", "tf32x1", "bf16"]
    _ffi.init(0)
    lib = _ffi.load_library()
    # queries: noisy copies of gallery rows
    rng = np.random.default_rng(0)
    pick = torch.from_numpy(rng.integers(0, N, size=Q)).cuda()
    qbase = torch.empty((Q, D), device="cuda")
    noise = torch.empty((Q, D), device="cuda")
    _ffi.check(lib.dif_synth_fill(_ffi.ptr(qbase), 3, 0, _ffi.ptr(pick), Q, D, None))
    _ffi.check(lib.dif_synth_fill(_ffi.ptr(noise), 33, 0, None, Q, D, None))
    torch.cuda.synchronize()
    q = qbase + 0.3 * noise
    flops = 2.0 * Q * N * D
    ref_rows = None
    out = []
    for mode in modes:
        for ctas in (2, 1):
            g = Gallery(N, D, "cosine", mode)
            g.set_option("gemm_ctas", ctas)
            g.set_option("l2_prefetch", int(os.environ.get("DIF_L2_PREFETCH", "0")))
            g.fill_synthetic(3, 0, N)
            for _ in range(2):
                s, ids = g.search(q, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 5
            kms = []
            e0.record()
            for _ in range(iters):
                s, ids = g.search(q, k)
            e1.record()
            torch.cuda.synchronize()
            kms.append(g.last_kernel_ms())
            ms = e0.elapsed_time(e1) / iters
            st = g.last_stats()
            top1 = float((ids[:, 0] == pick).float().mean())
            rows = ids.cpu().numpy()
            if ref_rows is None:
                ref_rows = rows
            same = bool((rows == ref_rows).all())
            rec = {"mode": mode, "ctas": ctas, "ms": round(ms, 3), "gemm_ms": round(kms[-1], 3),
                   "qps": round(Q / ms * 1e3), "gemm_tflops": round(flops / kms[-1] / 1e9, 1), "top1": top1,
                   "same_as_first": same, **st}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            g.close()
    # exactness on a sample against the oracle (CPU, canonical arithmetic)
    ns = 16
    t0 = time.time()
    gal = orc.synth_rows(3, 0, N, D)
    qs = q[:ns].cpu().numpy()
    ws, wr = orc.gallery_search(gal, qs, k, 1)
    dt = time.time() - t0
    ok = bool((wr == ref_rows[:ns]).all())
    print(json.dumps({"oracle_sample": ns, "oracle_s": round(dt, 2), "ids_match": ok, "threads": orc.num_threads()}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
