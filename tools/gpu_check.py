"""Staged bring-up checks for a fresh B200 box.  Each stage runs in its own process under a timeout so a
trap / hang in one kernel variant does not hide the others.  Summary -> gpurun_out/gpu_check.json.

    python tools/gpu_check.py            # all stages
    python tools/gpu_check.py gemm 2 1   # one stage in-process (precision 2, ctas 1)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def stage_simt():
    import numpy as np
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.gallery import Gallery
    from oracle import c_oracle as orc

    _ffi.init(0)
    lib = _ffi.load_library()
    D = 96
    out = torch.empty((1000, D), device="cuda")
    _ffi.check(lib.dif_synth_fill(_ffi.ptr(out), 7, 123, None, 1000, D, None))
    torch.cuda.synchronize()
    ref = orc.synth_rows(7, 123, 1000, D)
    assert np.array_equal(out.cpu().numpy(), ref), "synth mismatch"
    g = Gallery(1000, D, "cosine", "tf32x1")
    g.fill_synthetic(7, 123, 1000)
    got = g.rows()
    want = orc.normalize_rows(ref)
    assert np.array_equal(got, want), f"normalised rows differ: max {np.abs(got - want).max()}"
    print("simt ok")


def stage_gemm(prec: int, ctas: int):
    import numpy as np
    import torch

    from deep_insight_face_b200 import _ffi

    _ffi.init(0)
    lib = _ffi.load_library()
    res = {}
    for (M, N, K, splits) in [(128, 256, 64, 1), (300, 1000, 96, 1), (512, 4096, 512, 3), (257, 70000, 128, 37)]:
        torch.manual_seed(M + N)
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        Cd = torch.full((M, N), float("nan"), device="cuda")
        _ffi.check(lib.dif_debug_nt_gemm(_ffi.ptr(A), _ffi.ptr(B), M, N, K, _ffi.ptr(Cd), prec, ctas, splits, None))
        torch.cuda.synchronize()
        ref = A.double() @ B.double().T
        err = (Cd.double() - ref).abs().max().item()
        scale = (A.double().abs() @ B.double().abs().T).max().item()
        rel = err / scale
        tol = {0: 2e-6, 1: 8e-3, 2: 2e-3}[prec]
        res[f"{M}x{N}x{K}/{splits}"] = rel
        print(f"gemm prec={prec} ctas={ctas} {M}x{N}x{K} splits={splits}: max err {err:.3e} rel {rel:.3e}", flush=True)
        assert rel == rel and rel < tol, f"gemm prec={prec} ctas={ctas} {M}x{N}x{K}: rel err {rel} (tol {tol})"
    print("gemm ok", json.dumps(res))


def stage_search(prec: int, ctas: int, metric: int, resident: int = -1):
    import numpy as np
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.gallery import Gallery
    from oracle import c_oracle as orc

    for (N, Q, D, k) in [(1000, 37, 64, 10), (20000, 300, 128, 10), (70001, 513, 512, 5)]:
        rows = orc.synth_rows(3, 0, N, D)
        rng = np.random.default_rng(N)
        pick = rng.integers(0, N, size=Q)
        q = rows[pick] + 0.3 * orc.synth_rows(33, 0, Q, D)
        g = Gallery(N, D, "cosine" if metric == 1 else "l2", prec)
        g.set_option("gemm_ctas", ctas)
        g.set_option("resident_queries", resident)
        g.add(rows)
        s, ids, r = g.search(q, k, return_rows=True)
        st = g.last_stats()
        ws, wr = orc.gallery_search(rows, q, k, metric)
        bad = int((r.astype(np.int64) != wr).sum())
        sbad = int((s.view(np.uint32) != ws.view(np.uint32)).sum())
        top1 = float((r[:, 0] == pick).mean())
        print(f"search prec={prec} ctas={ctas} metric={metric} N={N} Q={Q} D={D} k={k}: row mismatches {bad}, "
              f"score-bit mismatches {sbad}, top1 recall {top1:.3f}, stats {st}", flush=True)
        assert bad == 0 and sbad == 0
        assert np.array_equal(ids, wr)
        # device-pointer entry point gives the same answer
        qd = torch.from_numpy(q).cuda()
        s2, ids2 = g.search(qd, k)
        assert np.array_equal(ids2.cpu().numpy(), wr)
        g.close()
    print("search ok")


STAGES = {
    "simt": lambda a: stage_simt(),
    "gemm": lambda a: stage_gemm(int(a[0]), int(a[1])),
    "search": lambda a: stage_search(*[int(x) for x in a]),
}


def main():
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]](sys.argv[2:])
        return
    os.makedirs(OUT, exist_ok=True)
    plan = [["simt"]]
    for prec, ctas in ((2, 1), (1, 2), (0, 2), (1, 18), (2, 18)):
        plan.append(["gemm", str(prec), str(ctas)])
    for prec, ctas, metric, res in ((0, 2, 1, -1), (1, 2, 1, -1), (2, 2, 1, -1), (1, 2, 0, -1), (2, 1, 1, -1),
                                    (1, 2, 1, 0), (0, 2, 0, -1)):
        plan.append(["search", str(prec), str(ctas), str(metric), str(res)])
    summary = {}
    for st in plan:
        name = "_".join(st)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), *st], capture_output=True, text=True,
                               timeout=240)
            ok, tail = r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        except subprocess.TimeoutExpired as e:
            ok, tail = False, "TIMEOUT " + str(e.stdout)[-1000:]
        summary[name] = {"ok": ok, "s": round(time.time() - t0, 1)}
        with open(os.path.join(OUT, f"check_{name}.log"), "w") as f:
            f.write(tail)
        print(f"[{'ok' if ok else 'FAIL'}] {name} ({summary[name]['s']} s)", flush=True)
        if not ok:
            print(tail[-1500:], flush=True)
    with open(os.path.join(OUT, "gpu_check.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
