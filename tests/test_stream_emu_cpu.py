"""The product's batched (streaming) path on the CPU: batch-wise noise samples and threshold recurrence, temporal3's
per-thread code, act4 / act, dst_sparse / dst_dense and the PPHT kernels -- the device code of csrc/*.cuh, emulated and
issued batch after batch like submit_impl / stream_kernel_launch / launch_hough_kernels issue it -- against golden
trajectories of the LIVE reference (MetLib/Detector.py:186-392): thresholds exact, snr to 1e-12, masks bit-exact, raw
Hough segments identical.  Three window shapes, dynamic mask on and off, both act kernels, ragged last batches.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_det_case, ragged_get
from emu_build import build

_SENS = {"low": 0, "normal": 1, "high": 2}


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    so = build(tmp_path_factory.mktemp("stream_emu"), "stream_path_emu.cpp",
               patched=["temporal3_kernel.cuh", "kernels_basic.cuh", "spatial_kernel.cuh", "hough.cuh"], shared=True)
    lib = C.CDLL(so)
    lib.emu_stream_path.restype = C.c_int
    return lib


@pytest.mark.parametrize("name,frames,batch", [("synth_320x240_n5_dyoff", 1000, 8), ("synth_384x216_n12_dyon_mask", 1000, 7),
                                               ("clip_192x144_n25", 1000, 16), ("synth_256x160_n6_fixed3_dense", 12, 5)])
def test_streaming_path_kernels_reproduce_the_reference_golden(emu_lib, name, frames, batch):
    if "dense" in name and not os.environ.get("EMU_SLOW"):
        pytest.skip("dense masks (tiers 2 / 3, > 500 segments per frame) take minutes under the emulator: set EMU_SLOW=1 "
                    "(last run: passed in 337 s)")
    g = load_det_case(name)
    T = min(frames, len(g["frames"]))
    fr = np.ascontiguousarray(g["frames"][:T])
    H, W = fr.shape[1:]
    c = g["cfg"]
    roi = (C.c_int * 4)(*[int(v) for v in g["std_roi"]])
    thr = np.zeros(T, np.int32); snr = np.zeros(T)
    dst = np.zeros((T, H, W), np.uint8); n_on = np.zeros(T, np.int32); nl = np.zeros(T, np.int32)
    raw = np.zeros((T, 512, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu_lib.emu_stream_path(p(fr), T, W, H, int(g["n"]), batch, int(c["adaptive"]), int(c["init_value"]), _SENS[c["sensitivity"]],
                                 int(c["interval"]), roi, *[int(v) for v in c["hough"]], int(c["dy_mask"]),
                                 C.c_double(float(g["mask_area"])), p(thr), p(snr), p(dst), p(n_on), p(nl), p(raw))
    assert rc == 0, rc
    assert np.array_equal(thr, g["bi_threshold"][:T])
    assert np.allclose(snr, g["snr"][:T], rtol=1e-12, atol=0)
    with_lines = 0
    for t in range(T):
        assert np.array_equal(dst[t], g["dst"][t]), (t, int(np.count_nonzero(dst[t] != g["dst"][t])))
        assert n_on[t] == np.count_nonzero(g["dst"][t])
        assert nl[t] == g["lines_num"][t], t
        if nl[t] <= 500:
            assert np.array_equal(raw[t, :nl[t]], ragged_get(g["raw_lines"], g["raw_offs"], t)), t
        with_lines += nl[t] > 0
    assert with_lines > 0
