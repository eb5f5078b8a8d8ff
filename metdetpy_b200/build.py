"""Builds libmetdet_b200.so in-tree with nvcc for sm_100a (no other architecture, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmetdet_b200.so")
SOURCES = ["metdet.cu"]
DEPS = ["metdet.cu", "common.cuh", "hough.cuh", "kernels_basic.cuh", "stream_kernel.cuh", "spatial_kernel.cuh", "temporal_kernel.cuh", "temporal3_kernel.cuh",
        "temporal3_dispatch.cuh", "preproc.cuh", "classic.cuh", "perframe_kernel.cuh", "mfnr.cuh",
        os.path.join("..", "..", "include", "metdet_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmetdet_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "--fmad=false",
           "-cudart", "static", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
