"""
Compile the CUDA sources of foldingdiff_b200 into an in-tree shared library (sm_100a only).

    python -m foldingdiff_b200._build        # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree to
the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libfoldingdiff_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "kernels_simt.cuh", "gemm_tc.cuh", "attention_mma.cuh", "attention_pool.cuh", "attention_tc.cuh", "writers.hpp", "nerf.cuh", "philox.cuh",
           os.path.join("..", "..", "include", "foldingdiff_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and os.path.getmtime(p) > built:
            return True
    return False


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH, "-lz"]
    if verbose:
        print("[foldingdiff_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)
