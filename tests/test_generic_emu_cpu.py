"""The product's generic per-frame path on the CPU: noise sample -> EMA / threshold recurrence -> fused_frame_kernel
(stack, diff, median, threshold, close, dynamic mask) -> PPHT kernels, the device code of csrc/kernels_basic.cuh and
csrc/hough.cuh run by the thread-block emulator and driven like mdb_update / mdb_detect drive it, against golden
trajectories of the LIVE reference (tests/golden/det_*.npz, MetLib/Detector.py:186-392): thresholds exact, snr to 1e-12,
masks bit-exact, raw Hough segments identical.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_det_case, ragged_get
from emu_build import build_generic

_SENS = {"low": 0, "normal": 1, "high": 2}


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    so = build_generic(tmp_path_factory.mktemp("generic_emu"))
    lib = C.CDLL(so)
    lib.emu_generic_path.restype = C.c_int
    return lib


@pytest.mark.parametrize("name,frames", [("synth_203x157_n3_high", 1000), ("synth_300x200_n7_low", 1000), ("clip_192x144_n25", 1000)])
def test_generic_path_kernels_reproduce_the_reference_golden(emu_lib, name, frames):
    g = load_det_case(name)
    T = min(frames, len(g["frames"]))
    fr = np.ascontiguousarray(g["frames"][:T])
    H, W = fr.shape[1:]
    mask = np.ascontiguousarray(g["mask"], np.uint8)
    c = g["cfg"]
    roi = (C.c_int * 4)(*[int(v) for v in g["std_roi"]])
    thr = np.zeros(T, np.int32); thrf = np.zeros(T); snr = np.zeros(T)
    dst = np.zeros((T, H, W), np.uint8); n_on = np.zeros(T, np.int32); nl = np.zeros(T, np.int32)
    raw = np.zeros((T, 512, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu_lib.emu_generic_path(p(fr), T, W, H, int(g["n"]), p(mask), 0, int(c["adaptive"]), int(c["init_value"]),
                                  _SENS[c["sensitivity"]], int(c["interval"]), roi, *[int(v) for v in c["hough"]], int(c["dy_mask"]),
                                  C.c_double(float(g["mask_area"])), p(thr), p(thrf), p(snr), p(dst), p(n_on), p(nl), p(raw))
    assert rc == 0, rc
    assert np.array_equal(thr, g["bi_threshold"][:T])
    assert np.allclose(snr, g["snr"][:T], rtol=1e-12, atol=0)
    assert np.allclose(thrf, g["bi_threshold_float"][:T], rtol=1e-12, atol=0)
    with_lines = 0
    for t in range(T):
        assert np.array_equal(dst[t], g["dst"][t]), (t, int(np.count_nonzero(dst[t] != g["dst"][t])))
        assert n_on[t] == np.count_nonzero(g["dst"][t])
        assert nl[t] == g["lines_num"][t], t
        want = ragged_get(g["raw_lines"], g["raw_offs"], t)
        if nl[t] <= 500:
            assert np.array_equal(raw[t, :nl[t]], want), t
        with_lines += nl[t] > 0
    assert with_lines > 0
