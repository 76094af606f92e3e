"""Build recipe for libmppi_b200.so (hand-written sm_100a CUDA behind a C ABI).

    python -m mppi_robotarm_b200.build          # or __graft_entry__.build()

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE next to this file so it
travels to the GPU box with the repository snapshot; it is git-ignored (source-only history).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmppi_b200.so")
SOURCES = [os.path.join(CSRC, "mppi_cabi.cu")]
DEPENDS = SOURCES + [os.path.join(CSRC, "mppi_kernels.cuh"), os.path.join(CSRC, "mppi_math.cuh"),
                     os.path.join(ROOT, "include", "mppi_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libmppi_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in DEPENDS)


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile the CUDA library for sm_100a; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
