"""Builds libspair_b200.so (the C-ABI kernel library) in-tree with nvcc for sm_100a.

The library has no torch / Python dependency: it is plain CUDA runtime code behind
``include/spair_b200.h``.  nvcc cross-compiles without a GPU, so this runs on the CPU-only
build container as well as on the B200 box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libspair_b200.so")
SOURCES = ["heads.cu", "glimpse.cu", "render.cu", "kl.cu", "sweep.cu", "stem.cu"]
HEADERS = ["common.cuh", "warp_math.cuh", os.path.join("..", "..", "include", "spair_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libspair_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("SPAIR_NVCC_EXTRA", "").split()      # e.g. -DSW_TIMING for tools/sweep_phase_timing.py
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libspair_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
