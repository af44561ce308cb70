"""
The C ABI from plain C: `include/foldingdiff_b200.h` must be valid strict C99 (not only C++),
the in-tree library must link against a C program with nothing but `-lfoldingdiff_b200`, and on a machine without a
GPU the program must fail LOUDLY at fd_create (no CPU fallback).  `examples/sample_host.c` is the program.
"""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from foldingdiff_b200 import _build

CSRC = os.path.dirname(_build.LIB_PATH)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if not os.path.isfile(_build.LIB_PATH):
        _build.build()
    out = str(tmp_path_factory.mktemp("cex") / "sample_host")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "sample_host.c"), "-L", CSRC, "-lfoldingdiff_b200", f"-Wl,-rpath,{CSRC}", "-lm", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_header_is_strict_c99_and_the_library_links_from_c(exe):
    assert os.access(exe, os.X_OK)


def test_c_program_fails_loudly_without_a_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure path cannot be reached")
    r = subprocess.run([exe, "4", "10"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1, (r.returncode, r.stdout, r.stderr)
    assert "ABI 2" in r.stdout
    assert "fd_create failed (2)" in r.stderr and "no CPU fallback" in r.stderr


def test_c_program_rejects_bad_arguments(exe):
    r = subprocess.run([exe, "0"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "usage" in r.stderr
