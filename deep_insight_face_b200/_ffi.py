"""ctypes binding of libdif_b200.so (the C ABI declared in include/dif_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is usable, every
compute entry point raises.  Nothing in this package imports anything from `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libdif_b200.so")

METRIC_SQL2 = 0
METRIC_COSINE = 1
PREC_TF32X3 = 0
PREC_BF16 = 1
PREC_TF32X1 = 2
PREC_BF16X3 = 3
MAX_TOPK = 24
LOSS_BH_COSINE = 0
LOSS_BH_EUCLIDEAN = 1
LOSS_SOFT_MARGIN = 4
TRANSPORT_NCCL = 0
TRANSPORT_PEER = 1
NCCL_ID_BYTES = 128

_PRECISIONS = {"tf32x3": PREC_TF32X3, "fp32": PREC_TF32X3, "bf16": PREC_BF16, "tf32": PREC_TF32X1,
               "tf32x1": PREC_TF32X1, "bf16x3": PREC_BF16X3}
_METRICS = {"cosine": METRIC_COSINE, "cos": METRIC_COSINE, "l2": METRIC_SQL2, "sql2": METRIC_SQL2,
            "euclidean": METRIC_SQL2}


class DifError(RuntimeError):
    """An entry point of libdif_b200.so returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libdif_b200 error {code}: {message}")
        self.code = code


def precision_code(p) -> int:
    if isinstance(p, str):
        try:
            return _PRECISIONS[p.lower()]
        except KeyError:
            raise ValueError(f"unknown precision {p!r}; expected one of {sorted(_PRECISIONS)}") from None
    return int(p)


def metric_code(m) -> int:
    if isinstance(m, str):
        try:
            return _METRICS[m.lower()]
        except KeyError:
            raise ValueError(f"unknown metric {m!r}; expected one of {sorted(_METRICS)}") from None
    return int(m)


_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); mirrors include/dif_b200.h one to one
SIGNATURES = {
    "dif_init": (_i32, [_i32]),
    "dif_last_error": (C.c_char_p, []),
    "dif_sync": (_i32, [_vp]),
    "dif_version": (C.c_char_p, []),
    "dif_launch_count": (_i64, []),
    "dif_release_retired": (_i64, []),
    "dif_gallery_create": (_vp, [_i32, _i64, _i32, _i32, _i32]),
    "dif_gallery_destroy": (None, [_vp]),
    "dif_gallery_add": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "dif_gallery_add_host": (_i32, [_vp, _vp, _vp, _i64]),
    "dif_gallery_fill_synth": (_i32, [_vp, _u64, _i64, _i64, _vp]),
    "dif_gallery_set_id_base": (_i32, [_vp, _i64]),
    "dif_gallery_remove": (_i32, [_vp, _vp, _i64, _vp]),
    "dif_gallery_get_ids": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "dif_gallery_size": (_i64, [_vp]),
    "dif_gallery_set_option": (_i32, [_vp, C.c_char_p, _i32]),
    "dif_gallery_reset": (_i32, [_vp]),
    "dif_gallery_search": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dif_gallery_search_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "dif_gallery_last_stats": (_i32, [_vp, C.POINTER(_i64)]),
    "dif_gallery_last_kernel_ms": (_i32, [_vp, C.POINTER(_f32)]),
    "dif_gallery_last_phase_ms": (_i32, [_vp, C.POINTER(_f32)]),
    "dif_gallery_get_rows": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "dif_topk_merge": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dif_nccl_unique_id": (_i32, [_vp]),
    "dif_nccl_comm_create": (_i32, [_i32, _i32, _vp, C.POINTER(_vp)]),
    "dif_nccl_comm_destroy": (_i32, [_vp]),
    "dif_shard_chunk_bytes": (_i64, [_i32, _i32, _i32]),
    "dif_gallery_search_packed": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dif_shard_merge": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dif_gallery_shard_attach": (_i32, [_vp, _vp, _i32, _i32, _i64, _i32, _i32, _i32]),
    "dif_gallery_search_sharded": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dif_gallery_search_sharded_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "dif_synth_fill": (_i32, [_vp, _u64, _i64, _vp, _i64, _i32, _vp]),
    "dif_debug_nt_gemm": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp]),
    "dif_debug_gemm_layout": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dif_debug_gemm_time": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(_f32)]),
    "dif_batch_hard": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "dif_batch_hard_host": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i32]),
    "dif_batch_hard_host_buffers": (_i32, [_i32, _i32, C.POINTER(_vp)]),
    "dif_batch_hard_set_path": (_i32, [_i32]),
    "dif_batch_all": (_i32, [_vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "dif_tfa_triplet": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _f32, _vp, _vp]),
    "dif_labels_from_onehot": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "dif_l2_normalize": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "dif_l2_normalize_bwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "dif_triplet_apn": (_i32, [_vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "dif_euclidean_distance": (_i32, [_vp, _vp, _i32, _i32, _f32, _vp, _vp]),
    "dif_contrastive_loss": (_i32, [_vp, _vp, _i32, _f32, _vp, _vp, _vp]),
    "dif_arcface": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "dif_arcface_host": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _i32]),
    "dif_pair_distance": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "dif_pair_distance_host": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "dif_fold_mean": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "dif_threshold_sweep": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "dif_threshold_sweep_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp]),
}

_lib = None
_lib_lock = threading.Lock()
_inited_devices: set[int] = set()


def load_library() -> C.CDLL:
    """dlopen libdif_b200.so and type every entry point.  Raises if the library is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  deep_insight_face_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing and not os.environ.get("DIF_ALLOW_PARTIAL_LIB"):
            raise ImportError(f"{LIB_PATH} does not export {missing}: header and library are out of sync; rebuild")
        _lib = lib
        return lib


def last_error() -> str:
    return load_library().dif_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise DifError(rc, last_error())


def init(device=None) -> None:
    """Select `device`, verify it is sm_100 and create its context (idempotent).  The library serves ONE device
    per process (dif_init refuses a second one): `device=None` means the device this process is already bound
    to, or device 0 for a fresh process - what the host-array entry points use."""
    if device is None:
        device = next(iter(_inited_devices)) if _inited_devices else 0
    device = int(device)
    if device in _inited_devices:   # dif_init queries device properties (milliseconds): once per device is enough
        return
    lib = load_library()
    check(lib.dif_init(device))
    _inited_devices.add(device)


def launch_count() -> int:
    return int(load_library().dif_launch_count())


def release_retired() -> int:
    """Free workspace blocks the loss entry points outgrew (kept alive for CUDA graphs captured at smaller sizes);
    call when no such graph is alive, e.g. after a sweep over batch sizes.  Returns the number of blocks freed."""
    return int(load_library().dif_release_retired())


_c_char = C.c_char


def ptr(a) -> int | None:
    """Raw address of a numpy array (host) or torch tensor (host or device); None passes NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        try:   # a.ctypes / __array_interface__ build helper objects (~2 us each, 7 pointers per loss call)
            return C.addressof(_c_char.from_buffer(a))
        except (TypeError, ValueError, BufferError):   # read-only, empty or non-contiguous
            return a.ctypes.data
    return int(a.data_ptr())


def host_array(a, dtype, shape=None) -> np.ndarray:
    """C-contiguous numpy view/copy of `a` with `dtype` (numpy in, torch-CPU in, or TF eager tensor in)."""
    if not isinstance(a, np.ndarray):
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        elif hasattr(a, "numpy"):
            a = a.numpy()
        else:
            a = np.asarray(a)
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def is_device_tensor(a) -> bool:
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


def current_stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream (the library orders its kernels on the caller's stream)."""
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)
