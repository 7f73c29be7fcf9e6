"""The C-ABI boundary: the library loads, exports every symbol include/dif_b200.h declares, the ctypes table
covers them all, and compute entry points fail loudly (never fall back) without a usable device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dif_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dif_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    fns = header_functions()
    for must in ("dif_gallery_search", "dif_gallery_search_host", "dif_batch_hard", "dif_arcface", "dif_pair_distance",
                 "dif_threshold_sweep", "dif_topk_merge", "dif_triplet_apn"):
        assert must in fns


def test_library_exports_every_declared_symbol(lib):
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, f"libdif_b200.so does not export {missing}"


def test_ctypes_table_matches_header(lib):
    from deep_insight_face_b200 import _ffi

    assert sorted(_ffi.SIGNATURES) == header_functions()


def test_no_device_fails_loudly(lib):
    import torch

    from deep_insight_face_b200 import _ffi

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    rc = lib.dif_init(0)
    assert rc == -3  # DIF_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.dif_last_error()
    with pytest.raises(_ffi.DifError):
        from deep_insight_face_b200.gallery import Gallery

        Gallery(16, 64)


def test_invalid_arguments_return_codes(lib):
    assert lib.dif_gallery_size(None) == -1
    assert lib.dif_topk_merge(None, None, None, 1, 1, 1, 1, None, None, None, None) == -1
    assert b"null" in lib.dif_last_error()
    assert lib.dif_pair_distance(None, None, 0, 0, 0, None, None, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep_insight_face_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|from\s+\.\.?oracle|dif_oracle\.so|libdif_oracle", txt, re.M):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, f"product code references the oracle: {offenders}"


def test_header_is_plain_c():
    """The boundary is a C ABI: include/dif_b200.h must compile as C99 (no C++ types, no torch types)."""
    import os
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest

        pytest.skip("gcc not available")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dif_b200.h")
    res = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
