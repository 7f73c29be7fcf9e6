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


def shard_range(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`: the first n_rows % world ranks hold one extra row."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedGallery:
    """Gallery rows partitioned contiguously over the ranks of a torch.distributed process group.

    Each rank searches its shard (local top-k with global row ids), one all-gather exchanges the
    k candidates per query per rank (Q*k*(4+8+8) bytes per rank) and every rank merges them with
    the same ordering key (score best-first, global row ascending), so the result equals the
    single-GPU search of the concatenated gallery bit for bit.
    """

    def __init__(self, n_rows_global: int, dim: int, metric="cosine", precision="tf32x3", device: int = 0,
                 group=None):
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_rows_global = int(n_rows_global)
        self.row_lo, self.row_hi = shard_range(n_rows_global, self.rank, self.world)
        self.metric = _ffi.metric_code(metric)
        self.local = Gallery(max(1, self.row_hi - self.row_lo), dim, metric, precision, device)
        self.local.set_id_base(self.row_lo)
        self.device = device
        self._explicit_ids = False

    def fill_synthetic(self, seed: int) -> None:
        self.local.fill_synthetic(seed, self.row_lo, self.row_hi - self.row_lo)

    def add_local(self, rows, ids=None) -> None:
        self._explicit_ids = self._explicit_ids or ids is not None
        self.local.add(rows, ids)

    def search(self, queries, k: int = 10):
        """queries: torch-CUDA [Q, dim], identical on every rank.  Returns (scores, ids, global_rows) tensors."""
        import torch

        scores, ids, rows = self.local.search(queries, k, return_rows=True)
        if not self._explicit_ids:
            # ids = id_base + row with id_base = row_lo: the id IS the global row (-1 in empty slots), so two
            # all-gathers (scores, ids) carry everything the merge needs
            if self.world == 1:
                return scores, ids, ids
            g_scores, g_ids = exchange_candidates(scores, ids, group=self.group)
            return merge_candidates(g_scores, g_ids, g_ids, self.metric)
        grows = torch.where(rows >= 0, rows.to(torch.int64) + self.row_lo, torch.full_like(ids, -1))
        if self.world == 1:
            return scores, ids, grows
        g_scores, g_ids, g_rows = exchange_candidates(scores, ids, grows, group=self.group)
        return merge_candidates(g_scores, g_rows, g_ids, self.metric)


    def search_host(self, queries, k: int = 10):
        """Host-facing call: numpy queries in, numpy (scores, ids) out.  One rank: the C-ABI host entry point
        (pinned staging, H2D, search, D2H).  Several ranks: pinned H2D here, then search/exchange/merge, D2H."""
        if self.world == 1:
            return self.local.search(queries, k)
        import torch

        q = _ffi.host_array(queries, np.float32)
        if getattr(self, "_pin", None) is None or self._pin.shape != q.shape:
            self._pin = torch.empty(q.shape, dtype=torch.float32).pin_memory()
        self._pin.numpy()[...] = q
        qd = self._pin.to(f"cuda:{self.device}", non_blocking=True)
        scores, ids, _ = self.search(qd, k)
        return scores.cpu().numpy(), ids.cpu().numpy()

    def close(self) -> None:
        self.local.close()


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
