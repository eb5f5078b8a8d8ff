"""Two of the SURVEY 8(f) rows on the CPU, device code under the thread-block emulator against golden vectors of the LIVE
reference: ClassicDetector (MetLib/Detector.py:245-299 -- csrc/classic.cuh kernels + noise / threshold kernels + PPHT with the
configured maxLineGap: thresholds, snr, masks, every segment) and the loader's Transform chain (MetLib/imgproc.py:82-139 --
csrc/preproc.cuh with the tap tables of mdb_preproc_axis_taps: resize, gray, mask, exposure merge, bit-exact).  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ragged_get
from emu_build import build_classic

_SENS = {"low": 0, "normal": 1, "high": 2}


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    so = build_classic(tmp_path_factory.mktemp("classic_emu"))
    lib = C.CDLL(so)
    lib.emu_classic_path.restype = C.c_int
    lib.emu_preproc.restype = C.c_int
    return lib


@pytest.mark.parametrize("name,batch,frames", [("synth_320x240", 5, 1000), ("odd_203x157_mask", 16, 24), ("dense_256x160_fixed", 4, 1000)])
def test_classic_detector_kernels_reproduce_the_reference_golden(emu_lib, name, batch, frames):
    from metdetpy_b200.detector import select_subarea
    g = np.load(os.path.join(GOLDEN, f"classic_{name}.npz"))
    fr = np.ascontiguousarray(g["frames"][:frames])
    T, H, W = fr.shape
    want_dst = np.unpackbits(g["dst_bits"][:T], axis=1)[:, :H * W].reshape(T, H, W) * np.uint8(255)
    adaptive, init_value, area, interval = g["cfg"]
    mask = np.ascontiguousarray(g["mask"], np.uint8)
    roi = (C.c_int * 4)(*[int(v) for v in select_subarea(mask, float(area))])
    cap = 8192
    thr = np.zeros(T, np.int32); snr = np.zeros(T); dst = np.zeros((T, H, W), np.uint8); nl = np.zeros(T, np.int32)
    raw = np.zeros((T, cap, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu_lib.emu_classic_path(p(fr), T, W, H, batch, int(adaptive), int(init_value), _SENS[str(g["sens"])], int(interval), roi,
                                  *[int(v) for v in g["hough"]], C.c_double(float(mask.sum())), p(thr), p(snr), p(dst), p(nl), p(raw), cap)
    assert rc == 0, rc
    assert np.array_equal(thr, g["bi_threshold"][:T])
    assert np.allclose(snr, g["snr"][:T], rtol=1e-12, atol=0)
    total = 0
    for t in range(T):
        assert np.array_equal(dst[t], want_dst[t]), (t, int(np.count_nonzero(dst[t] != want_dst[t])))
        ref = ragged_get(g["raw_lines"], g["raw_offs"], t)
        if t >= 3:  # the first three frames return no lines (Detector.py:264-265); their masks are empty
            assert nl[t] == len(ref) and np.array_equal(raw[t, :nl[t]], ref), t
        total += len(ref)
    assert total > 0


def test_loader_preprocessing_kernel_reproduces_the_reference_golden(emu_lib):
    from metdetpy_b200 import _lib
    lib = _lib.load()
    g = np.load(os.path.join(GOLDEN, "preproc.npz"))
    for name in g["names"]:
        W, H, gray, exp = (int(v) for v in g[f"{name}_cfg"])
        fr = np.ascontiguousarray(g[f"{name}_frames"])
        T, H0, W0 = fr.shape[:3]
        ch = 3 if fr.ndim == 4 else 1
        mask = np.ascontiguousarray(g[f"{name}_mask"], np.uint8)
        taps = []
        for dst_n, src_n, clamp, stride in ((W, W0, 1, ch), (H, H0, 0, 1)):
            a = [np.zeros(dst_n, np.int32) for _ in range(4)]
            assert lib.mdb_preproc_axis_taps(dst_n, src_n, clamp, *[x.ctypes.data for x in a]) == 0
            a[0] *= stride; a[1] *= stride
            taps.append(np.ascontiguousarray(np.stack(a, 1)))
        G = (T + exp - 1) // exp
        out = np.zeros((G, H, W), np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = emu_lib.emu_preproc(p(fr), T, W0, H0, ch, 0, W, H, int((W0, H0) != (W, H)), exp, p(taps[0]), p(taps[1]), p(mask), p(out))
        assert rc == 0
        assert np.array_equal(out, g[f"{name}_out"]), name
