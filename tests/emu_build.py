"""Build helper of the CPU emulation tests: compiles a harness under tests/emu/ with g++ against the product's kernel
sources.  Kernels that use dynamic shared memory or PTX hints get two mechanical edits in a scratch copy (`extern __shared__`
-> `extern`; `asm volatile("prefetch...")` lines dropped); common.cuh is copied with its header include made path-independent."""
import os
import re
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "metdetpy_b200", "csrc")


def cuda_include():
    for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include")):
        if os.path.exists(os.path.join(p, "cuda_runtime.h")):
            return p
    return None


def patched_sources(tmp, names):
    """Scratch copies `<name>` -> `<stem>_emu.cuh` of kernel headers (+ common.cuh), edited for the block emulator."""
    for name in names:
        src = open(os.path.join(CSRC, name)).read()
        src = src.replace("extern __shared__", "extern")
        src = re.sub(r"\n[^\n]*asm volatile\(\"prefetch\.global\.L2[^\n]*", "\n", src)
        # host-side launchers (<<< >>> syntax) inside kernel headers are not part of the device code under test
        src = re.sub(r"static inline void launch_noise_samples\(.*?\n}\n", "", src, flags=re.S)
        if name == "temporal_kernel.cuh":  # temporal2's inline-PTX helpers -> tests/emu/t2_ptx_shims.h
            def cut(text, first, stop):
                a, b = text.index(first), text.index(stop)
                assert a < b, (first, stop)
                return text[:a] + text[b:]
            src = cut(src, "__device__ __forceinline__ void t2_commit()", "// bytes 1 and 3 of x as clean u16x2 lanes")
            src = cut(src, "template <int WPT>\n__device__ __forceinline__ void t2_cp(", "// per-thread running state of the window")
            src = cut(src, "// base + idx * stride as one IMAD.WIDE", "// Compute stage.")
            src = src.replace("#define T2_K 8", '#define T2_K 8\n#include "t2_ptx_shims.h"', 1)
            assert "asm" not in src.split("#pragma once", 1)[1], "temporal_kernel.cuh has PTX the shims do not cover"
        for other in names:  # headers under test that include each other pick up the scratch copies
            src = src.replace(f'#include "{other}"', f'#include "{other.replace(".cuh", "_emu.cuh")}"')
        open(os.path.join(tmp, name.replace(".cuh", "_emu.cuh")), "w").write(src)
    # the default window -> shape table of temporal3_dispatch.cuh as `case n: temporal3_batch<U, BL, P, K>(...)` lines, so that
    # an emulated path runs a window with exactly the shape the product launches for it
    disp = open(os.path.join(CSRC, "temporal3_dispatch.cuh")).read()
    table = disp[disp.rindex("switch (n) {"):]
    table = table[:table.index("default: return -2;")]
    rows = re.findall(r"T3_CASE\((\d+), (\d+), (\d+), (\d+), (\d+), \d+\)", table)
    assert len(rows) >= 27, len(rows)
    with open(os.path.join(tmp, "t3_shapes_emu.inc"), "w") as f:
        for n, u, bl, pp, k in rows:
            f.write(f"case {n}: temporal3_batch<{u}, {bl}, {pp}, {k}>(T3_SHAPE_ARGS); break;\n")
    c = open(os.path.join(CSRC, "common.cuh")).read().replace('#include "../../include/metdet_b200.h"', '#include "metdet_b200.h"')
    open(os.path.join(tmp, "common.cuh"), "w").write(c)


_BUILT = {}  # one build per harness and process: several test modules share the larger libraries


def build(tmp, harness, patched=(), extra_c=(), std="c++20", opt="-O1", shared=False):
    key = (harness, tuple(patched), tuple(extra_c), std, opt, shared)
    if key in _BUILT and os.path.exists(_BUILT[key]):
        return _BUILT[key]
    _BUILT[key] = _build(tmp, harness, patched, extra_c, std, opt, shared)
    return _BUILT[key]


def _build(tmp, harness, patched, extra_c, std, opt, shared):
    if shutil.which("g++") is None or shutil.which("gcc") is None:
        pytest.skip("no g++ / gcc")
    inc = cuda_include()
    if inc is None:
        pytest.skip("CUDA headers not found")
    tmp = str(tmp)
    patched_sources(tmp, patched)
    objs = []
    for c in extra_c:
        o = os.path.join(tmp, os.path.basename(c) + ".o")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-c", os.path.join(REPO, c), "-o", o])
        objs.append(o)
    exe = os.path.join(tmp, os.path.splitext(os.path.basename(harness))[0] + (".so" if shared else ""))
    subprocess.check_call(["g++", opt, f"-std={std}", "-ffp-contract=off"] + (["-shared", "-fPIC"] if shared else []) + ["-I", tmp, "-I", os.path.join(REPO, "tests", "emu"),
                           "-I", os.path.join(REPO, "include"), "-I", inc, os.path.join(REPO, "tests", "emu", harness)] + objs +
                          ["-o", exe])
    return exe


# the three larger harness libraries, shared by several test modules and tests/emu_fuzz.py
STREAM_PATCHED = ["temporal3_kernel.cuh", "temporal_kernel.cuh", "kernels_basic.cuh", "spatial_kernel.cuh", "hough.cuh", "perframe_kernel.cuh"]
GENERIC_PATCHED = ["kernels_basic.cuh", "hough.cuh"]
CLASSIC_PATCHED = ["kernels_basic.cuh", "spatial_kernel.cuh", "classic.cuh", "preproc.cuh", "hough.cuh"]


def build_stream(tmp):
    return build(tmp, "stream_path_emu.cpp", patched=STREAM_PATCHED, shared=True)


def build_generic(tmp):
    return build(tmp, "generic_path_emu.cpp", patched=GENERIC_PATCHED, shared=True)


def build_classic(tmp):
    return build(tmp, "classic_path_emu.cpp", patched=CLASSIC_PATCHED, shared=True)
