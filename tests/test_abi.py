"""The C-ABI shared library: loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import torch

from conftest import ROOT
from foldingdiff_b200 import _build, _native


def header_functions():
    text = open(os.path.join(ROOT, "include", "foldingdiff_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.isfile(_build.LIB_PATH), "run __graft_entry__.build()"
    assert os.path.dirname(_build.LIB_PATH).startswith(ROOT)


def test_exports_every_declared_symbol():
    lib = _native.lib()
    declared = header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_native.SIGNATURES), set(declared) ^ set(_native.SIGNATURES)


def test_version_and_counts():
    lib = _native.lib()
    assert lib.fd_abi_version() == 2
    assert lib.fd_num_weights(12) == 4 + 17 * 12 + 6
    assert b"sm_100a" in lib.fd_build_info()
    assert lib.fd_last_error() is not None


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        return
    lib = _native.lib()
    dims = _native.FdDims(192, 1, 6, 384, 128, 6, 1, 1e-12, 1e-12)
    n = lib.fd_num_weights(1)
    bufs = [torch.zeros(192 * 384) for _ in range(n)]
    arr = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
    tt, coef = torch.zeros(192), torch.zeros(4)
    h = C.c_void_p()
    rc = lib.fd_create(C.byref(dims), arr, n, tt.data_ptr(), coef.data_ptr(), 0, 0, C.byref(h))
    assert rc == 2  # FD_ERR_CUDA
    assert b"no CPU fallback" in lib.fd_last_error() or b"CUDA" in lib.fd_last_error()
    assert lib.fd_set_batch(None, 1, 1, None, 0, None, None) == 1  # FD_ERR_INVALID on null handle


def test_library_staleness_is_decided_by_source_hash(tmp_path):
    """build() reuses the in-tree .so only when it was built from exactly the sources in the tree (content hash recorded
    next to it), whatever the mtimes say - a stale library must never be what the GPU tests load."""
    assert os.path.isfile(_build.SHA_PATH), "built by an older _build: run python -m foldingdiff_b200._build"
    assert not _build._stale()
    recorded = open(_build.SHA_PATH).read()
    try:
        os.utime(os.path.join(_build.CSRC, "api.cu"))          # newer mtime, same content: still fresh
        assert not _build._stale()
        with open(_build.SHA_PATH, "w") as f:                    # a library built from other sources: stale
            f.write("0" * 64 + "\n")
        assert _build._stale()
    finally:
        with open(_build.SHA_PATH, "w") as f:
            f.write(recorded)
    assert not _build._stale()
