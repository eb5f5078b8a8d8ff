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
from emu_build import build_stream

_SENS = {"low": 0, "normal": 1, "high": 2}


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    lib = C.CDLL(build_stream(tmp_path_factory.mktemp("stream_emu")))
    lib.emu_stream_path.restype = C.c_int
    return lib


@pytest.mark.parametrize("name,frames,batch", [("synth_320x240_n5_dyoff", 1000, 8), ("synth_384x216_n12_dyon_mask", 1000, 7),
                                               ("clip_192x144_n25", 1000, 16), ("synth_256x160_n6_fixed3_dense", 16, 5),
                                               ("clip_cfg1_480x270_n6", 1000, 16), ("clip_cfg1_960x540_n6_range", 1000, 22)])
def test_streaming_path_kernels_reproduce_the_reference_golden(emu_lib, name, frames, batch):
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
    assert emu_lib.emu_noise16_launches() > 0  # the noise kernel the product selects for aligned frames (16-byte loads)


@pytest.mark.parametrize("name,batch", [("synth_384x216_n12_dyon_mask", 7), ("clip_192x144_n25", 16)])
def test_streaming_path_with_the_second_generation_temporal_kernel(emu_lib, name, batch):
    """temporal_version = 2 (temporal2_kernel, the kernel windows without a temporal3 shape get) against the same goldens."""
    emu_lib.emu_set_temporal_version(2)
    try:
        test_streaming_path_kernels_reproduce_the_reference_golden(emu_lib, name, 1000, batch)
        assert emu_lib.emu_temporal2_launches() > 0
    finally:
        emu_lib.emu_set_temporal_version(3)


@pytest.mark.parametrize("name,world,batch", [("synth_384x216_n12_dyon_mask", 3, 7), ("clip_192x144_n25", 2, 16)])
def test_time_sharded_protocol_on_the_cpu(emu_lib, name, world, batch):
    """SURVEY 8(e) with product code only: every virtual rank computes the integer noise sums of its chunk's sample timers
    (noise_sample_kernel, emulated), the sums are pooled (the all-gather), the threshold recurrence is replayed by the
    library's host code (mdb_replay_thresholds), and every rank runs seek + (2n-2)-frame halo + its chunk with the replayed
    thresholds through the emulated streaming kernels: thresholds, masks and raw segments of every frame equal the golden
    trajectory of the live (sequential) reference."""
    from metdetpy_b200 import sharding as S
    g = load_det_case(name)
    fr = np.ascontiguousarray(g["frames"])
    T, H, W = fr.shape
    n, c = int(g["n"]), g["cfg"]
    roi_t = [int(v) for v in g["std_roi"]]
    roi = (C.c_int * 4)(*roi_t)
    roi_px = (roi_t[2] - roi_t[0]) * (roi_t[3] - roi_t[1])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    shards = S.plan_shards(T, world, n)
    samples = []
    for sh in shards:
        taus = np.ascontiguousarray(S.sample_timers(sh.start, sh.end, n, int(c["interval"])), np.int64)
        if len(taus) == 0:
            continue
        lo = max(0, sh.start - (n - 1))
        part = np.ascontiguousarray(fr[lo:sh.end])
        sums = np.zeros((len(taus), 2), np.uint64)
        assert emu_lib.emu_noise_sums(p(part), len(part), C.c_longlong(lo), W, H, n, int(c["interval"]), roi, p(taus), len(taus), p(sums)) == 0
        samples += [(int(t), int(a), int(b)) for t, (a, b) in zip(taus, sums)]
    thr, thr_f, snr = S.replay_thresholds_native(samples, roi_px, n, 0, T, adaptive=c["adaptive"], init_value=c["init_value"],
                                                 sensitivity=c["sensitivity"], interval=c["interval"])
    assert np.array_equal(thr, g["bi_threshold"]) and np.allclose(snr, g["snr"], rtol=1e-12, atol=0)
    for sh in shards:
        part = np.ascontiguousarray(fr[sh.halo_start:sh.end])
        m = len(part)
        halo = sh.start - sh.halo_start
        assert halo == min(sh.start, 2 * n - 2)
        thr_in = np.ascontiguousarray(thr[sh.halo_start:sh.end], np.int32)
        o_thr = np.zeros(m, np.int32); o_snr = np.zeros(m); dst = np.zeros((m, H, W), np.uint8)
        n_on = np.zeros(m, np.int32); nl = np.zeros(m, np.int32); raw = np.zeros((m, 512, 4), np.int32)
        rc = emu_lib.emu_stream_chunk(p(part), m, C.c_longlong(sh.halo_start), halo, p(thr_in), W, H, n, batch, roi,
                                      *[int(v) for v in c["hough"]], int(c["dy_mask"]), C.c_double(float(g["mask_area"])),
                                      p(o_thr), p(o_snr), p(dst), p(n_on), p(nl), p(raw))
        assert rc == 0, rc
        for t in range(sh.start, sh.end):
            k = t - sh.halo_start
            assert np.array_equal(dst[k], g["dst"][t]), (sh.rank, t, int(np.count_nonzero(dst[k] != g["dst"][t])))
            assert nl[k] == g["lines_num"][t], (sh.rank, t)
            if nl[k] <= 500:
                assert np.array_equal(raw[k, :nl[k]], ragged_get(g["raw_lines"], g["raw_offs"], t)), (sh.rank, t)


@pytest.mark.parametrize("name,frames", [("synth_320x240_n5_dyoff", 1000), ("synth_384x216_n12_dyon_mask", 1000), ("clip_192x144_n25", 1000),
                                         ("synth_256x160_n6_fixed3_dense", 14), ("clip_cfg1_480x270_n6", 1000)])
def test_per_frame_resident_state_path_reproduces_the_reference_golden(emu_lib, name, frames):
    """update(); detect() frame by frame on the O(1) path (pf_update() / mdb_detect() with bits_ready in csrc/metdet.cu):
    staging copy, noise sample + threshold, pf_update_kernel in two halves, suffix rebuild at block ends, act / dst / PPHT on
    the predicate bits it leaves behind -- against the golden trajectory of the live reference."""
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
    rc = emu_lib.emu_perframe_path(p(fr), T, W, H, int(g["n"]), int(c["adaptive"]), int(c["init_value"]), _SENS[c["sensitivity"]],
                                   int(c["interval"]), roi, *[int(v) for v in c["hough"]], int(c["dy_mask"]),
                                   C.c_double(float(g["mask_area"])), p(thr), p(snr), p(dst), p(n_on), p(nl), p(raw))
    assert rc == 0, rc
    assert np.array_equal(thr, g["bi_threshold"][:T])
    assert np.allclose(snr, g["snr"][:T], rtol=1e-12, atol=0)
    for t in range(T):
        assert np.array_equal(dst[t], g["dst"][t]), (t, int(np.count_nonzero(dst[t] != g["dst"][t])))
        assert nl[t] == g["lines_num"][t], t
        if nl[t] <= 500:
            assert np.array_equal(raw[t, :nl[t]], ragged_get(g["raw_lines"], g["raw_offs"], t)), t
