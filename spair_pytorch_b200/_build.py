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
SOURCES = ["heads.cu", "glimpse.cu", "render.cu", "kl.cu", "sweep.cu", "stem.cu", "gemm.cu", "conv.cu"]
HEADERS = ["common.cuh", "warp_math.cuh", "sweep_tc.cuh", os.path.join("..", "..", "include", "spair_b200.h")]
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart", "shared"]


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
    """Compiles every csrc/*.cu to an object (one nvcc process per source, in parallel), then links the shared library.
    Objects whose source and headers are older than the object are reused unless ``force``."""
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("SPAIR_NVCC_EXTRA", "").split()      # e.g. -DSW_TIMING for tools/sweep_phase_timing.py
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_time = max(hdr_time, os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(obj_dir, s[:-3] + ".o")
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time)
        if force or extra or not fresh:
            jobs.append(([nvcc] + COMPILE_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj], s))
    procs = [(src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for cmd, src in jobs]
    failed = []
    for src, proc in procs:
        log = proc.communicate()[0]
        if proc.returncode != 0:
            sys.stderr.write(log)
            failed.append(src)
        elif verbose:
            sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc failed compiling %s" % ", ".join(failed))
    res = subprocess.run([nvcc] + LINK_FLAGS + objs + ["-o", LIB_PATH], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libspair_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
