import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """libdif_b200.so, built here if nvcc is present and the .so is missing (the GPU box receives it prebuilt)."""
    from deep_insight_face_b200 import _ffi

    if not os.path.exists(_ffi.LIB_PATH):
        from deep_insight_face_b200 import build

        build.build_library()
    return _ffi.load_library()


@pytest.fixture(scope="session")
def orc():
    from oracle import c_oracle

    c_oracle.lib()
    return c_oracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "verification_reference.npz"))


@pytest.fixture(scope="session")
def gpu(lib):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deep_insight_face_b200 import _ffi

    _ffi.init(0)
    return torch.device("cuda", 0)
