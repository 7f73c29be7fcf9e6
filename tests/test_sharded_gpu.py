"""The row-sharded search behind the C ABI (csrc/shard.cu: dif_gallery_search_packed, dif_shard_merge,
dif_gallery_shard_attach, dif_gallery_search_sharded[_host]) against the CPU oracle.

 * ranks emulated on ONE GPU: per-shard packed chunks + dif_shard_merge == the single-gallery oracle;
 * a one-rank NCCL communicator made by the library (dif_nccl_unique_id / dif_nccl_comm_create): both
   transports, device and host entry points;
 * two real ranks (two processes, two GPUs, NCCL) when the box has two devices - skipped on a one-GPU box.
The reference has no distributed code; the contract is "equal to the one-gallery search".
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make(orc, N, Q, D, seed=3):
    rows = orc.synth_rows(seed, 0, N, D)
    pick = np.random.default_rng(N + Q).integers(0, N, size=Q)
    q = rows[pick] + 0.3 * orc.synth_rows(seed + 30, 0, Q, D)
    return rows, q, pick


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("explicit_ids", [False, True])
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_packed_chunks_and_shard_merge(gpu, orc, world, explicit_ids, metric):
    import torch

    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.gallery import Gallery, shard_range

    lib = _ffi.load_library()
    N, Q, D, k = 10007, 120, 128, 10
    rows, q, _ = make(orc, N, Q, D)
    ids_all = np.random.default_rng(5).permutation(N).astype(np.int64) + 7_000_000
    cb = lib.dif_shard_chunk_bytes(Q, k, int(explicit_ids))
    assert cb == (Q * k * (16 if explicit_ids else 8) + 15) // 16 * 16
    chunks = torch.zeros(world * cb, dtype=torch.uint8, device=gpu)
    info = np.zeros((world, 2), dtype=np.int64)
    qd = torch.from_numpy(q).cuda()
    for rank in range(world):
        lo, hi = shard_range(N, rank, world)
        with Gallery(hi - lo, D, metric, "bf16x3") as g:
            g.set_id_base(1000 + lo)              # default id = 1000 + global row: ids and rows are told apart
            g.add(rows[lo:hi], ids_all[lo:hi] if explicit_ids else None)
            _ffi.check(lib.dif_gallery_search_packed(g._h, _ffi.ptr(qd), Q, k, int(explicit_ids),
                                                     chunks.data_ptr() + rank * cb, None))
            if explicit_ids:   # an id-less chunk of a gallery with explicit ids would lose them: refused
                assert lib.dif_gallery_search_packed(g._h, _ffi.ptr(qd), Q, k, 0, chunks.data_ptr() + rank * cb, None) == -5
            torch.cuda.synchronize()
        info[rank] = (lo, 1000 + lo)
    info_d = torch.from_numpy(info).cuda()
    s = torch.empty((Q, k), dtype=torch.float32, device=gpu)
    ids = torch.empty((Q, k), dtype=torch.int64, device=gpu)
    gr = torch.empty((Q, k), dtype=torch.int64, device=gpu)
    mcode = _ffi.metric_code(metric)
    _ffi.check(lib.dif_shard_merge(chunks.data_ptr(), _ffi.ptr(info_d), world, Q, k, mcode, int(explicit_ids),
                                   _ffi.ptr(s), _ffi.ptr(ids), _ffi.ptr(gr), None))
    torch.cuda.synchronize()
    ws, wr = orc.gallery_search(rows, q, k, mcode)
    assert np.array_equal(gr.cpu().numpy(), wr)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), ws.view(np.uint32))
    assert np.array_equal(ids.cpu().numpy(), ids_all[wr] if explicit_ids else wr + 1000)


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_sharded_search_on_a_one_rank_communicator(gpu, orc, transport):
    """world = 1 through the full C path: library-made NCCL communicator, attach, device and host entry points,
    growing (Q, k) re-attaches, a short gallery leaves -1 slots."""
    import torch

    from deep_insight_face_b200.gallery import ShardedGallery

    N, Q, D, k = 6001, 300, 128, 10
    rows, q, _ = make(orc, N, Q, D)
    g = ShardedGallery(N, D, "cosine", "tf32x3", device=0, transport=transport, single_rank_exchange=True)
    try:
        g.add_local(rows)
        ws, wr = orc.gallery_search(rows, q, k, 1)
        for _ in range(3):   # epochs advance, both parities of the peer buffers are used
            s, ids, gr = g.search(torch.from_numpy(q).cuda(), k)
            assert np.array_equal(gr.cpu().numpy(), wr) and np.array_equal(ids.cpu().numpy(), wr)
            assert np.array_equal(s.cpu().numpy().view(np.uint32), ws.view(np.uint32))
        assert g.transport == transport and g._own_comm
        hs, hi = g.search_host(q, k)
        assert np.array_equal(hi, wr) and np.array_equal(hs.view(np.uint32), ws.view(np.uint32))
        pinned = torch.from_numpy(q).pin_memory().numpy()
        hs, hi = g.search_host(pinned, k, bcast_root=0)
        assert np.array_equal(hi, wr) and np.array_equal(hs.view(np.uint32), ws.view(np.uint32))
        assert g.search_host(q, k, want_result=False) is None
        hs, hi = g.search_host(q, k, bcast_root=-2)      # sliced upload degenerates to a full upload on one rank
        assert np.array_equal(hi, wr) and np.array_equal(hs.view(np.uint32), ws.view(np.uint32))
        # more queries and a larger k than the buffers were attached for
        q2 = np.concatenate([q, q[:50] * 2.0])
        ws2, wr2 = orc.gallery_search(rows, q2, 24, 1)
        hs, hi = g.search_host(q2, 24)
        assert np.array_equal(hi, wr2) and np.array_equal(hs.view(np.uint32), ws2.view(np.uint32))
    finally:
        g.close()
    g = ShardedGallery(5, 64, "l2", "bf16", device=0, transport=transport, single_rank_exchange=True)
    try:
        few = orc.synth_rows(9, 0, 5, 64)
        g.add_local(few, np.arange(5, dtype=np.int64) * 3 + 11)
        s, ids, gr = g.search(torch.from_numpy(few[:2]).cuda(), 8)
        ws, wr = orc.gallery_search(few, few[:2], 8, 0)
        assert np.array_equal(gr.cpu().numpy(), wr)
        assert np.array_equal(ids.cpu().numpy(), np.where(wr >= 0, wr * 3 + 11, -1))
    finally:
        g.close()


def test_sharded_errors(gpu, orc, lib):
    from deep_insight_face_b200 import _ffi
    from deep_insight_face_b200.gallery import Gallery

    with Gallery(100, 64) as g:
        out = (C.c_float * 10)()
        # not attached
        assert lib.dif_gallery_search_sharded(g._h, 1, 0, 1, 1, 1, 1, out, out, None, None) == -5
        assert b"dif_gallery_shard_attach" in lib.dif_last_error()
        assert lib.dif_gallery_shard_attach(g._h, None, 0, 1, 0, 16, 10, 0) == -1
    assert lib.dif_shard_chunk_bytes(0, 10, 0) == -1 and lib.dif_shard_chunk_bytes(3, 25, 0) == -1
    assert lib.dif_nccl_comm_create(2, 5, None, None) == -1
    assert _ffi.TRANSPORT_PEER == 1


def _two_rank_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import torch
    import torch.distributed as dist

    from deep_insight_face_b200.gallery import ShardedGallery, shard_range
    from oracle import c_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    notes = []
    try:
        N, Q, D, k = 40009, 333, 128, 10
        rows, q, _ = make(orc, N, Q, D)
        ids_all = np.random.default_rng(5).permutation(N).astype(np.int64) + 7_000_000
        ws, wr = orc.gallery_search(rows, q, k, 1)
        lo, hi = shard_range(N, rank, world)
        for transport in ("peer", "nccl"):
            for explicit in (False, True):
                for own in (False, True):
                    if own:
                        os.environ["DIF_OWN_NCCL_COMM"] = "1"
                    else:
                        os.environ.pop("DIF_OWN_NCCL_COMM", None)
                    g = ShardedGallery(N, D, "cosine", "bf16x3", device=rank, transport=transport)
                    g.add_local(rows[lo:hi], ids_all[lo:hi] if explicit else None)
                    want_ids = ids_all[wr] if explicit else wr
                    for step in range(4):
                        s, ids, gr = g.search(torch.from_numpy(q).cuda(), k)
                        good = (np.array_equal(gr.cpu().numpy(), wr) and np.array_equal(ids.cpu().numpy(), want_ids)
                                and np.array_equal(s.cpu().numpy().view(np.uint32), ws.view(np.uint32)))
                        ok = ok and good
                    for root in (-2, -1, 0, 1):
                        res = g.search_host(q if (root < 0 or root == rank) else np.zeros_like(q), k, bcast_root=root,
                                            want_result=(rank == 0))
                        if rank == 0:
                            ok = ok and np.array_equal(res[1], want_ids) and np.array_equal(res[0].view(np.uint32), ws.view(np.uint32))
                        else:
                            ok = ok and res is None
                    notes.append(f"{transport}->{g.transport} ids={explicit} own_comm={g._own_comm}")
                    g.close()
    except Exception as e:  # noqa: BLE001 - reported through the result file
        ok = False
        notes.append(repr(e))
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(("ok" if ok else "mismatch") + "\n" + "\n".join(notes))
    dist.destroy_process_group()


def test_two_ranks_two_gpus(gpu, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        txt = open(tmp_path / f"rank{r}.txt").read()
        assert txt.startswith("ok"), txt
