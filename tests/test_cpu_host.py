"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/metdet_b200.h
declares, host-side logic (NMS in the library's C++, ROI selection, EMA, config marshalling)
matches the reference's golden vectors, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, REPO, assert_nms_equivalent, has_len2_ties, ragged_get
from metdetpy_b200 import BinaryCfg, _lib
from metdetpy_b200 import detector as D


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "metdet_b200.h")).read()
    declared = set(re.findall(r"\b(mdb_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib._build.build())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared <= set(_lib.SYMBOLS), declared - set(_lib.SYMBOLS)


def test_struct_layouts_match_header(tmp_path):
    assert ctypes.sizeof(_lib.Config) == 4 * (7 + 4 + 7 + 4)
    assert ctypes.sizeof(_lib.FrameInfo) == 8 + 4 + 4 + 8 * 4 + 4 * 4
    # the header is plain C: compile it with gcc and compare every field offset with the ctypes mirror
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"mdb_config": _lib.Config, "mdb_frame_info": _lib.FrameInfo, "mdb_mfnr_params": _lib.MfnrParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "metdet_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
    lib = _lib.load()
    assert lib.mdb_version() >= 100
    assert lib.mdb_device_count() >= 0


def test_nms_host_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    exact = 0
    for k in range(len(g["lines_offs"]) - 1):
        lines = ragged_get(g["lines"], g["lines_offs"], k)
        ref = ragged_get(g["out"], g["out_offs"], k)
        refp = ragged_get(g["prob"], g["out_offs"], k)
        out, p = D.lineset_nms(lines)
        assert_nms_equivalent(out, p, ref, refp, lines, k)
        exact += not has_len2_ties(lines)
    assert exact >= 20
    out, p = D.lineset_nms(np.zeros((0, 4), np.int32))
    assert out.shape == (0, 4) and p.shape == (0,)


def test_select_subarea_and_ema_golden():
    g = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    for k in range(4):
        H, W = (int(v) for v in g["roi_mask_shapes"][k])
        m = np.unpackbits(g[f"roi_mask{k}"])[:H * W].reshape(H, W)
        assert D.select_subarea(m, float(g["roi_areas"][k])) == tuple(int(v) for v in g["rois"][k])
    e = D.EMA(float(g["ema_momentum"]), float(g["ema_warmup"]))
    for v, r in zip(g["ema_in"], g["ema_out"]):
        e.update(v)
        assert e.cur_value == r
    with pytest.raises(ValueError):
        D.select_subarea(np.ones((10, 10), np.uint8), 0)


def test_argument_errors_without_gpu():
    lib = _lib.load()
    assert lib.mdb_destroy(None) != 0
    assert b"null" in lib.mdb_last_error()
    cfg = _lib.Config()
    h = ctypes.c_void_p()
    m = np.ones((4, 4), np.uint8)
    assert lib.mdb_create(ctypes.byref(cfg), m.ctypes.data, ctypes.byref(h)) == -1  # 0x0 frame
    cfg.width = cfg.height = 4
    cfg.window = 5000
    assert lib.mdb_create(ctypes.byref(cfg), m.ctypes.data, ctypes.byref(h)) == -1  # window > MDB_MAX_WINDOW
    assert b"window" in lib.mdb_last_error()
    k = ctypes.c_int32()
    assert lib.mdb_lineset_nms(None, -1, None, None, ctypes.byref(k)) == -1


def test_no_cpu_fallback():
    if _lib.load().mdb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.MetDetError, match="no CPU fallback"):
        D.M3Detector(1, 10, np.ones((32, 32), np.uint8), 10, BinaryCfg())
    from metdetpy_b200 import stacker
    with pytest.raises(_lib.MetDetError):
        stacker.merge_max(np.zeros((2, 8, 8), np.uint8))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) use it."""
    for top in ("metdetpy_b200", "scripts", "examples", "include"):
        for root, _, files in os.walk(os.path.join(REPO, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".sh")):
                    src = open(os.path.join(root, f), errors="replace").read()
                    assert not re.search(r"^\s*(from|import)\s+oracle|__import__\(\"oracle|oracle/|liboracle", src, re.M), \
                        f"{top}/{f} reaches into oracle/"


def test_batched_tie_renms_equals_the_checker_frame_by_frame():
    """mdb_lineset_nms_frames with numpy's own argsort order per frame == the CPU checker's lineset_nms (which calls
    np.argsort like the reference, utils.py:804) on tie-heavy segment sets of every size up to 40."""
    import ctypes as C
    from metdetpy_b200 import _lib
    from oracle import m3_oracle as O
    lib = _lib.load()
    rng = np.random.default_rng(9)
    T, M = 64, _lib.MAX_LINES
    infos = (_lib.FrameInfo * T)()
    raw = np.zeros((T, M, 4), np.int32)
    lines = np.zeros((T, M, 4), np.int32)
    prob = np.zeros((T, M), np.float64)
    for i in range(T):
        n = 1 + i % 40
        p0 = rng.integers(0, 60, (n, 2))
        d = rng.integers(-6, 7, (n, 2)) * 3          # few distinct lengths: many ties
        raw[i, :n] = np.concatenate([p0, p0 + d], 1)
        infos[i].n_raw = n
    fr = np.arange(T, dtype=np.int32)
    off = np.concatenate(([0], np.cumsum([infos[i].n_raw for i in range(T)]))).astype(np.int64)
    orders = np.empty(off[-1], np.int32)
    for i in range(T):
        seg = raw[i, :infos[i].n_raw]
        l2 = np.power(seg[:, 3] - seg[:, 1], 2) + np.power(seg[:, 2] - seg[:, 0], 2)
        orders[off[i]:off[i + 1]] = np.argsort(l2)[::-1]
    rc = lib.mdb_lineset_nms_frames(T, fr.ctypes.data, orders.ctypes.data, off.ctypes.data, raw.ctypes.data, C.byref(infos),
                                    lines.ctypes.data, prob.ctypes.data)
    assert rc == 0, lib.mdb_last_error()
    for i in range(T):
        ref, rp = O.lineset_nms(raw[i, :infos[i].n_raw].copy())
        k = infos[i].n_lines
        assert k == len(ref) and np.array_equal(lines[i, :k], np.asarray(ref).reshape(-1, 4)), i
        assert np.allclose(prob[i, :k], rp, rtol=1e-12, atol=0, equal_nan=True), i
    bad = orders.copy(); bad[off[5]] = bad[off[5] + 1]
    assert lib.mdb_lineset_nms_frames(T, fr.ctypes.data, bad.ctypes.data, off.ctypes.data, raw.ctypes.data, C.byref(infos),
                                      lines.ctypes.data, prob.ctypes.data) != 0


def test_c_caller_links_against_the_library(tmp_path):
    """examples/c_abi_example.c: a plain C program compiled against include/metdet_b200.h and linked with the shared
    library (what a non-Python host of the boundary does).  Without a GPU it must report that and exit 0."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib = _lib._build.build()
    exe = tmp_path / "c_abi_example"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
                           os.path.join(REPO, "examples", "c_abi_example.c"), "-L", os.path.dirname(lib), "-lmetdet_b200",
                           "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    assert f"ABI version {_lib.ABI_VERSION}" in out
    if _lib.load().mdb_device_count() == 0:
        assert "no CPU fallback" in out
    else:
        assert "lines, first" in out  # the moving bar is found


def test_every_kernel_of_the_product_runs_under_the_cpu_emulator():
    """Every __global__ function under metdetpy_b200/csrc is launched by at least one harness under tests/emu/ (which the CPU
    tests run against golden vectors / the checker); temporal3_kernel through its per-thread body t3::thread_main, which the
    __global__ wrapper calls once per thread after filling the per-frame table.  Exceptions are listed with the reason."""
    import glob
    not_emulated = {
        "t3_table_kernel": "per-frame table of the bulk-copy / L2-prefetch feeds of temporal3 (tuning variants, not a default shape)",
    }
    kernels = set()
    for f in glob.glob(os.path.join(REPO, "metdetpy_b200", "csrc", "*.cu*")):
        src = open(f).read()
        kernels |= set(re.findall(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(", src))
    assert len(kernels) >= 35, sorted(kernels)
    harness = "".join(open(f).read() for f in glob.glob(os.path.join(REPO, "tests", "emu", "*.cpp")))
    missing = sorted(k for k in kernels if k not in not_emulated and not re.search(r"\b%s\b" % k, harness))
    assert not missing, missing
