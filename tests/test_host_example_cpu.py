"""examples/host_search.c - a plain-C99 host on include/dif_b200.h - builds with -Wall -Wextra -Werror, links against
libdif_b200.so and, in this GPU-less container, fails loudly with the library's "no CPU fallback" message.  Its own
logic (argument order, buffer sizes, the double-precision checks it makes and their tolerances) is then exercised
against tests/fake/fake_dif_b200.c, a CPU stand-in built into a temp dir whose arithmetic is the canonical oracle's.
On a B200 the same program runs against the real library (tests/test_zz5_host_example_gpu.py)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "host_search.c")
CFLAGS = ["-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include")]


def build_example(out, libdir):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    cmd = [gcc, *CFLAGS, SRC, "-L" + libdir, "-ldif_b200", "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined", "-lm", "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_example_links_against_the_library_and_fails_loudly_without_a_gpu(lib, tmp_path):
    import torch

    exe = build_example(str(tmp_path / "host_search"), os.path.join(ROOT, "deep_insight_face_b200"))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is tests/test_zz5_host_example_gpu.py")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 3 and "no CPU fallback" in res.stderr, (res.returncode, res.stderr)


def test_example_logic_against_the_cpu_stand_in(orc, tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    oracle_dir = os.path.join(ROOT, "oracle")
    fake = str(tmp_path / "libdif_b200.so")
    cmd = [gcc, "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "fake", "fake_dif_b200.c"), "-L" + oracle_dir, "-ldif_oracle",
           "-Wl,-rpath," + oracle_dir, "-lm", "-o", fake]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    exe = build_example(str(tmp_path / "host_search_fake"), str(tmp_path))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "host_search OK" in res.stdout, (res.returncode, res.stdout, res.stderr)
