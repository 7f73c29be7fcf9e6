"""1:N gallery search: the identity database behind `verify` / `compare_faces`, as a device-resident index.

The reference keeps identities in a python dict and compares one pair per call
(deep_insight_face/predictions.py:104-150, api.py:94-104).  `Gallery` is the batched form named
by BASELINE.json: add embedding rows (+ identity ids), then `search(queries, k)` returns, per
query, the k best rows ordered by (score best-first, row index ascending).  Scores are cosine
similarity (descending) or squared L2 (ascending) in the canonical fp32 arithmetic of
csrc/dif_canon.cuh, so ids AND scores are bit-reproducible whatever tensor-core mode filtered them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi


class Gallery:
    """Device-resident embedding index on one B200.

    precision: "tf32x3" (fp32-exact tensor-core filter), "bf16x3", "bf16", "tf32" - see include/dif_b200.h.
    """

    def __init__(self, capacity: int, dim: int, metric="cosine", precision="tf32x3", device: int = 0):
        self._lib = _ffi.load_library()
        _ffi.init(device)
        self.device = int(device)
        self.dim = int(dim)
        self.capacity = int(capacity)
        self.metric = _ffi.metric_code(metric)
        self.precision = _ffi.precision_code(precision)
        self._h = self._lib.dif_gallery_create(self.device, self.capacity, self.dim, self.metric, self.precision)
        if not self._h:
            raise _ffi.DifError(-1, _ffi.last_error())

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.dif_gallery_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self) -> int:
        return int(self._lib.dif_gallery_size(self._h))

    # ------------------------------------------------------------------ building
    def add(self, rows, ids=None) -> None:
        """Append rows [n, dim] (numpy / torch-CPU -> staged H2D; torch-CUDA -> in place) and optional int64 ids."""
        if _ffi.is_device_tensor(rows):
            import torch

            rows = rows.contiguous().float()
            if rows.dim() != 2 or rows.shape[1] != self.dim:
                raise ValueError(f"rows must be [n, {self.dim}]")
            if ids is not None:
                ids = ids.to(device=rows.device, dtype=torch.int64).contiguous()
            st = _ffi.current_stream_ptr(rows.device)
            _ffi.check(self._lib.dif_gallery_add(self._h, _ffi.ptr(rows), _ffi.ptr(ids), rows.shape[0], st))
            torch.cuda.current_stream(rows.device).synchronize()  # rows/ids may be freed by the caller
            return
        rows = _ffi.host_array(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}]")
        if ids is not None:
            ids = _ffi.host_array(ids, np.int64, (rows.shape[0],))
        _ffi.check(self._lib.dif_gallery_add_host(self._h, _ffi.ptr(rows), _ffi.ptr(ids), rows.shape[0]))

    def fill_synthetic(self, seed: int, row0: int, n: int) -> None:
        """Append rows [row0, row0+n) of the shared counter-based synthetic gallery (generated on device)."""
        _ffi.check(self._lib.dif_gallery_fill_synth(self._h, int(seed), int(row0), int(n), None))
        _ffi.check(self._lib.dif_sync(None))

    def set_id_base(self, id_base: int) -> None:
        """ids of rows added without explicit ids are id_base + local row (a shard's global offset)."""
        _ffi.check(self._lib.dif_gallery_set_id_base(self._h, int(id_base)))

    def set_option(self, name: str, value: int) -> None:
        _ffi.check(self._lib.dif_gallery_set_option(self._h, name.encode(), int(value)))

    def reset(self) -> None:
        _ffi.check(self._lib.dif_gallery_reset(self._h))

    def ids(self, row0: int = 0, n: int | None = None) -> np.ndarray:
        """ids of the stored rows (explicit ids, or id_base + row) as a host int64 array."""
        import torch

        n = len(self) - row0 if n is None else n
        out = torch.empty(n, dtype=torch.int64, device=f"cuda:{self.device}")
        _ffi.check(self._lib.dif_gallery_get_ids(self._h, int(row0), int(n), _ffi.ptr(out),
                                                 _ffi.current_stream_ptr(out.device)))
        return out.cpu().numpy()

    def remove(self, ids=None, rows=None) -> int:
        """Delete identities by id (every row carrying one of `ids`) or by row number; the remaining rows keep
        their order and ids.  Returns the number of rows removed.  (The persistent index of SURVEY 8f row 4,
        replacing the python dict of predictions.py:112.)"""
        if (ids is None) == (rows is None):
            raise ValueError("pass exactly one of ids / rows")
        if ids is not None:
            rows = np.flatnonzero(np.isin(self.ids(), np.asarray(ids, dtype=np.int64).reshape(-1)))
        rows = np.unique(np.asarray(rows, dtype=np.int64).reshape(-1))
        if rows.size:
            _ffi.check(self._lib.dif_gallery_remove(self._h, _ffi.ptr(rows), int(rows.size), None))
        return int(rows.size)

    def rows(self, row0: int = 0, n: int | None = None) -> np.ndarray:
        """Canonical stored rows (normalised for cosine) as a host array."""
        import torch

        n = len(self) - row0 if n is None else n
        out = torch.empty((n, self.dim), dtype=torch.float32, device=f"cuda:{self.device}")
        _ffi.check(self._lib.dif_gallery_get_rows(self._h, int(row0), int(n), _ffi.ptr(out),
                                                  _ffi.current_stream_ptr(out.device)))
        return out.cpu().numpy()

    # ------------------------------------------------------------------ search
    def search(self, queries, k: int = 10, return_rows: bool = False):
        """Top-k per query.  numpy / torch-CPU queries take the host entry point (pinned staging, H2D,
        D2H, synchronises) and return numpy; torch-CUDA queries run stream-ordered and return tensors.

        Returns (scores [Q,k] fp32, ids [Q,k] int64[, rows [Q,k] int32]).
        """
        if _ffi.is_device_tensor(queries):
            import torch

            q = queries.contiguous().float()
            if q.dim() != 2 or q.shape[1] != self.dim:
                raise ValueError(f"queries must be [Q, {self.dim}]")
            Q = q.shape[0]
            scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
            ids = torch.empty((Q, k), dtype=torch.int64, device=q.device)
            rows = torch.empty((Q, k), dtype=torch.int32, device=q.device)
            if Q:
                _ffi.check(self._lib.dif_gallery_search(self._h, _ffi.ptr(q), Q, int(k), _ffi.ptr(scores),
                                                        _ffi.ptr(ids), _ffi.ptr(rows),
                                                        _ffi.current_stream_ptr(q.device)))
            return (scores, ids, rows) if return_rows else (scores, ids)
        q = _ffi.host_array(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}]")
        Q = q.shape[0]
        scores = np.empty((Q, k), dtype=np.float32)
        ids = np.empty((Q, k), dtype=np.int64)
        rows = np.empty((Q, k), dtype=np.int32)
        if Q:
            _ffi.check(self._lib.dif_gallery_search_host(self._h, _ffi.ptr(q), Q, int(k), _ffi.ptr(scores),
                                                         _ffi.ptr(ids), _ffi.ptr(rows)))
        return (scores, ids, rows) if return_rows else (scores, ids)

    def last_stats(self) -> dict:
        out = (C.c_int64 * 6)()
        _ffi.check(self._lib.dif_gallery_last_stats(self._h, out))
        return {"fallback_queries": int(out[0]), "kernels": int(out[1]), "splits": int(out[2]),
                "candidates_per_split": int(out[3]), "resident_queries": int(out[4])}

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _ffi.check(self._lib.dif_gallery_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)


    def last_phase_ms(self) -> dict:
        """Phase durations of the last search (synchronise the stream first)."""
        out = (C.c_float * 5)()
        _ffi.check(self._lib.dif_gallery_last_phase_ms(self._h, out))
        return {"prep": out[0], "filter": out[1], "rerank": out[2], "exact": out[3], "exchange_merge": out[4]}


def shard_range(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`: the first n_rows % world ranks hold one extra row."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedGallery:
    """Gallery rows partitioned contiguously over the ranks of a torch.distributed process group (one process per GPU).

    Each rank searches its shard, the k candidates per query per rank are exchanged as packed chunks (8 B per
    candidate) and every rank merges them with the same ordering key (score best-first, global row ascending), so
    the result equals the single-GPU search of the concatenated gallery bit for bit.  The whole step runs inside
    libdif_b200.so (`dif_gallery_search_sharded`, csrc/shard.cu) on the process group's NCCL communicator:

        transport="peer"  one kernel stores the chunk into every peer's buffer over NVLink (CUDA IPC), signals,
                          waits and merges - the exchange fused with its consumer (default; falls back to "nccl"
                          when the peers' buffers cannot be mapped, e.g. one visible device per process)
        transport="nccl"  one in-place ncclAllGather, then the merge kernel
    """

    def __init__(self, n_rows_global: int, dim: int, metric="cosine", precision="tf32x3", device: int = 0,
                 group=None, transport: str = "peer", single_rank_exchange: bool = False):
        import torch.distributed as dist

        self.group = group
        self._dist = dist.is_initialized()
        self.rank = dist.get_rank(group) if self._dist else 0
        self.world = dist.get_world_size(group) if self._dist else 1
        self.n_rows_global = int(n_rows_global)
        self.row_lo, self.row_hi = shard_range(n_rows_global, self.rank, self.world)
        self.metric = _ffi.metric_code(metric)
        self.local = Gallery(max(1, self.row_hi - self.row_lo), dim, metric, precision, device)
        self.local.set_id_base(self.row_lo)
        self.device = device
        self.dim = int(dim)
        if transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self.transport = transport
        self._lib = self.local._lib
        self._comm = None          # ncclComm_t as an int
        self._own_comm = False
        self._attached = None      # (max_queries, max_k) the exchange buffers were sized for
        self._explicit_ids = False
        # one rank has nobody to exchange with: search() is the plain local search unless a test asks for the full path
        self._direct = self.world == 1 and not single_rank_exchange

    # ------------------------------------------------------------------ building
    def fill_synthetic(self, seed: int) -> None:
        self.local.fill_synthetic(seed, self.row_lo, self.row_hi - self.row_lo)

    def add_local(self, rows, ids=None) -> None:
        if ids is not None:
            self._attached = None      # explicit ids change the packed chunk layout: exchange buffers are re-made
            self._explicit_ids = True
        self.local.add(rows, ids)

    # ------------------------------------------------------------------ communicator
    def _communicator(self) -> int:
        """ncclComm_t of the group: torch's own communicator when it can be borrowed (ProcessGroupNCCL._comm_ptr),
        else one created by the library from an id rank 0 broadcasts through the process group."""
        if self._comm is not None:
            return self._comm
        import ctypes as C

        import torch
        import torch.distributed as dist

        dev = torch.device("cuda", self.device)
        if self._dist and not os.environ.get("DIF_OWN_NCCL_COMM"):
            try:
                pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
                dist.all_reduce(torch.zeros(1, device=dev), group=self.group)   # the communicator exists after a collective
                torch.cuda.synchronize(dev)
                self._comm = int(pg._get_backend(dev)._comm_ptr())
                return self._comm
            except Exception:
                self._comm = None
        uid = (C.c_char * _ffi.NCCL_ID_BYTES)()
        if self.rank == 0:
            _ffi.check(self._lib.dif_nccl_unique_id(uid))
        if self._dist:
            box = [bytes(uid.raw)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                       group=self.group)
            uid = (C.c_char * _ffi.NCCL_ID_BYTES).from_buffer_copy(box[0])
        comm = C.c_void_p()
        _ffi.check(self._lib.dif_nccl_comm_create(self.world, self.rank, uid, C.byref(comm)))
        self._comm, self._own_comm = int(comm.value), True
        return self._comm

    def _attach(self, n_queries: int, k: int) -> None:
        if self._attached and n_queries <= self._attached[0] and k <= self._attached[1]:
            return
        mq = max(n_queries, self._attached[0] if self._attached else 0)
        mk = max(k, self._attached[1] if self._attached else 0)
        comm = self._communicator()
        code = _ffi.TRANSPORT_PEER if self.transport == "peer" else _ffi.TRANSPORT_NCCL
        rc = self._lib.dif_gallery_shard_attach(self.local._h, comm, self.rank, self.world, self.row_lo, mq, mk, code)
        if rc != 0 and code == _ffi.TRANSPORT_PEER and "DIF_TRANSPORT_PEER" in _ffi.last_error():
            # the peers' regions cannot be mapped (every rank gets this status together): exchange through NCCL
            self.transport = "nccl"
            rc = self._lib.dif_gallery_shard_attach(self.local._h, comm, self.rank, self.world, self.row_lo, mq, mk,
                                                    _ffi.TRANSPORT_NCCL)
        _ffi.check(rc)
        self._attached = (mq, mk)

    # ------------------------------------------------------------------ search
    def search(self, queries, k: int = 10):
        """queries: torch-CUDA [Q, dim], identical on every rank.  Returns (scores, ids, global_rows) tensors;
        stream-ordered on torch's current stream, no host synchronisation."""
        import torch

        q = queries.contiguous().float()
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}]")
        Q = q.shape[0]
        if self._direct:
            scores, ids, rows = self.local.search(q, k, return_rows=True)
            if not self._explicit_ids:      # default ids are row_lo + row: the id IS the global row
                return scores, ids, ids
            return scores, ids, torch.where(rows >= 0, rows.to(torch.int64) + self.row_lo, torch.full_like(ids, -1))
        self._attach(Q, k)
        scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
        ids = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        grows = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        _ffi.check(self._lib.dif_gallery_search_sharded(self.local._h, self._comm, self.rank, self.world, _ffi.ptr(q), Q,
                                                        int(k), _ffi.ptr(scores), _ffi.ptr(ids), _ffi.ptr(grows),
                                                        _ffi.current_stream_ptr(q.device)))
        return scores, ids, grows

    def search_host(self, queries, k: int = 10, bcast_root: int = -1, want_result: bool = True):
        """Host-facing call: numpy queries in, numpy (scores, ids) out, through dif_gallery_search_sharded_host.

        bcast_root = -1: every rank passes the (same) queries and uploads its own copy - straight from the caller's
        buffer when it is page-locked.  bcast_root = -2: every rank uploads 1/world of the rows, one NCCL all-gather
        over NVLink assembles the batch (the PCIe bytes per rank shrink by the world size).  bcast_root = r >= 0: only
        rank r's `queries` is read (the other ranks pass any array of the same shape): uploaded once, NCCL-broadcast.
        want_result=False skips this rank's D2H (returns None)."""
        i_upload = bcast_root < 0 or bcast_root == self.rank
        q = _ffi.host_array(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}]")
        Q = q.shape[0]
        if self._direct:
            res = self.local.search(q, k)
            return res if want_result else None
        self._attach(Q, k)
        scores = np.empty((Q, k), dtype=np.float32) if want_result else None
        ids = np.empty((Q, k), dtype=np.int64) if want_result else None
        _ffi.check(self._lib.dif_gallery_search_sharded_host(self.local._h, self._comm, self.rank, self.world,
                                                             _ffi.ptr(q) if i_upload else None, int(bcast_root), Q, int(k),
                                                             _ffi.ptr(scores), _ffi.ptr(ids), None))
        return (scores, ids) if want_result else None

    def close(self) -> None:
        self.local.close()        # frees the shard state (exchange buffers, IPC mappings) with the handle
        if self._own_comm and self._comm:
            self._lib.dif_nccl_comm_destroy(self._comm)
        self._comm = None


def exchange_candidates(*tensors, group=None):
    """All-gather the per-rank candidate lists: returns one [world, Q, k] tensor per input (NCCL and gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = []
    for t in tensors:
        # concatenation along dim 0 (the layout both NCCL and gloo accept), viewed as [world, Q, k]
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        out.append(buf.view((world,) + tuple(t.shape)))
    return tuple(out)


def merge_candidates(g_scores, g_rows, g_ids, metric):
    """Merge [world, Q, k] candidate lists on the device (csrc/gallery.cu:topk_merge_kernel)."""
    import torch

    lib = _ffi.load_library()
    world, Q, k = g_scores.shape
    scores = torch.empty((Q, k), dtype=torch.float32, device=g_scores.device)
    rows = torch.empty((Q, k), dtype=torch.int64, device=g_scores.device)
    ids = torch.empty((Q, k), dtype=torch.int64, device=g_scores.device)
    _ffi.check(lib.dif_topk_merge(_ffi.ptr(g_scores.contiguous()), _ffi.ptr(g_rows.contiguous()),
                                  _ffi.ptr(g_ids.contiguous()), world, Q, k, _ffi.metric_code(metric),
                                  _ffi.ptr(scores), _ffi.ptr(rows), _ffi.ptr(ids),
                                  _ffi.current_stream_ptr(g_scores.device)))
    return scores, ids, rows
