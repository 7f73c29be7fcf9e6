"""The N > 1 path on CPU: two gloo ranks run the product's shard layout and candidate exchange
(deep_insight_face_b200.gallery.shard_range / exchange_candidates); the per-shard search and the merge are
played by the oracle (there is no GPU here), and the result must equal the single-gallery search bit for bit."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from deep_insight_face_b200.gallery import exchange_candidates, shard_range
    from oracle import c_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, Q, D, k = 5003, 40, 64, 10
    rows = orc.synth_rows(3, 0, N, D)
    q = rows[:Q] + 0.3 * orc.synth_rows(33, 0, Q, D)
    lo, hi = shard_range(N, rank, world)
    s, r = orc.gallery_search(rows[lo:hi], q, k, 1)
    grow = np.where(r >= 0, r + lo, -1)
    g_s, g_i, g_r = exchange_candidates(torch.from_numpy(s), torch.from_numpy(grow), torch.from_numpy(grow))
    assert g_s.shape == (world, Q, k)
    ms, mr = orc.topk_merge(g_s.numpy(), g_r.numpy(), 1)
    ws, wr = orc.gallery_search(rows, q, k, 1)
    ok = np.array_equal(mr, wr) and np.array_equal(ms.view(np.uint32), ws.view(np.uint32))
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_exchange_and_merge(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_shard_ranges_partition_the_gallery():
    sys.path.insert(0, ROOT)
    from deep_insight_face_b200.gallery import shard_range

    for n, w in ((100_000_000, 8), (1_000_003, 4), (7, 8), (5, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
