"""examples/host_search.c - the plain-C host on include/dif_b200.h - built with gcc on the GPU box and run against the
real libdif_b200.so: 1:N search from host buffers, pair distances in both metrics and one batch-hard triplet step,
each checked inside the program against double-precision loops written out in C.  (Its logic runs against a CPU
stand-in in tests/test_host_example_cpu.py; written after the round's GPU budget was spent, first run is the driver's.)"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_runs_the_path(gpu, tmp_path):
    from test_host_example_cpu import build_example

    exe = build_example(str(tmp_path / "host_search"), os.path.join(ROOT, "deep_insight_face_b200"))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "host_search OK" in res.stdout, (res.returncode, res.stdout[-2000:], res.stderr[-2000:])
