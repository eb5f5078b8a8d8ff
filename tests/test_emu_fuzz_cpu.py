"""Randomised parity without a GPU (tests/emu_fuzz.py): the product's streaming and per-frame kernels under the CPU block
emulator against the CPU checker (cv2 backend = the reference's own calls, MetLib/Detector.py:186-392) on random small
configurations -- frame shapes, windows with every kind of temporal3 shape (whole ring in registers, half in the shared page,
sub-blocked), batch lengths that cut the van Herk blocks anywhere, adaptive / fixed thresholds, dynamic mask, Hough parameters,
polygon masks, flashes and flickering hot regions.  Thresholds, masks and raw Hough segments must be identical, snr to 1e-12.
A handful of fixed seeds here; `python tests/emu_fuzz.py FIRST COUNT` runs as many as wanted."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fuzz():
    pytest.importorskip("cv2")
    import emu_fuzz
    return emu_fuzz


@pytest.fixture(scope="module")
def lib(fuzz, tmp_path_factory):
    return fuzz.build_lib(str(tmp_path_factory.mktemp("emu_fuzz")))


@pytest.fixture(scope="module")
def generic_lib(fuzz, tmp_path_factory):
    return fuzz.build_generic_lib(str(tmp_path_factory.mktemp("emu_fuzz_generic")))


@pytest.mark.parametrize("seed", [2, 4, 10, 11, 18, 21, 22, 27, 36])
def test_random_configuration_equals_the_checker(fuzz, lib, seed):
    """Streaming path and per-frame resident-state path; seeds 2, 18, 21, 22, 36 have the mask applied inside the kernels' loads
    (apply_mask = 1) instead of by the loader."""
    case = fuzz.make_case(seed)
    for per_frame in (False, True):
        res = fuzz.run_case(lib, case, per_frame)
        assert res is None, (seed, per_frame, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, case["cfg"], res)


@pytest.mark.parametrize("seed", [7, 8, 15])
def test_windows_without_a_temporal3_shape_run_temporal2(fuzz, lib, seed):
    """n = 38, 13, 17: the product launches temporal2_kernel (csrc/temporal_kernel.cuh; its inline PTX replaced by
    tests/emu/t2_ptx_shims.h) -- what a 29.97 fps or 23.976 fps video with a one-second window gets."""
    case = fuzz.make_case(seed)
    before = lib.emu_temporal2_launches()
    res = fuzz.run_case(lib, case)
    assert res is None, (seed, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, case["cfg"], res)
    assert lib.emu_temporal2_launches() > before


@pytest.mark.parametrize("seed", [0, 2, 4, 9, 10, 11])
def test_temporal2_forced_equals_the_checker(fuzz, lib, seed):
    """temporal_version = 2 on windows that do have a temporal3 shape (incl. n = 60: sub-blocked van Herk, 4 blocks of 15)."""
    case = fuzz.make_case(seed)
    lib.emu_set_temporal_version(2)
    try:
        res = fuzz.run_case(lib, case)
    finally:
        lib.emu_set_temporal_version(3)
    assert res is None, (seed, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, case["cfg"], res)


@pytest.mark.parametrize("seed", [0, 2, 9])
def test_random_dense_configuration_equals_the_checker(fuzz, lib, seed):
    """Low fixed thresholds on a noisy sky: thousands of on-pixels per frame (word-list overflows -> dst_dense, PPHT tiers 1b / 2)."""
    case = fuzz.make_case(seed, dense=True)
    for per_frame in (False, True):
        res = fuzz.run_case(lib, case, per_frame)
        assert res is None, (seed, per_frame, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, case["cfg"], res)


@pytest.mark.parametrize("seed", [0, 1, 2, 7, 9, 11])
def test_random_configuration_through_the_generic_kernels(fuzz, generic_lib, seed):
    """Any width, any window (1, 31, 38, 129, 140 ...), mask on the device or applied by the loader."""
    case = fuzz.make_case(seed, any_width=True)
    res = fuzz.run_case(generic_lib, case, generic=True)
    assert res is None, (seed, {k: case[k] for k in ("W", "H", "n", "T", "apply_mask")}, case["cfg"], res)


@pytest.fixture(scope="module")
def classic_lib(fuzz, tmp_path_factory):
    return fuzz.build_classic_lib(str(tmp_path_factory.mktemp("emu_fuzz_classic")))


@pytest.mark.parametrize("seed", [1, 3, 6, 9, 12])
def test_random_configuration_through_the_classic_detector(fuzz, classic_lib, seed):
    case = fuzz.make_case(seed, any_width=True)
    res = fuzz.run_classic_case(classic_lib, case)
    assert res is None, (seed, {k: case[k] for k in ("W", "H", "T", "batch")}, case["cfg"], res)


def test_random_loader_transform_chains(fuzz, classic_lib):
    """Random source / target sizes (down- and up-scaling, no resize), colour and gray sources, exposure merges of 1..4 frames
    with a ragged last group, random masks."""
    for seed in range(40):
        res = fuzz.run_preproc_case(classic_lib, seed)
        assert res is None, (seed, res)


@pytest.mark.parametrize("seed", [2, 7, 11, 15, 18, 23])
def test_random_time_sharded_run_equals_the_sequential_run(fuzz, lib, seed):
    """SURVEY 8(e) on random cases with 2..5 virtual ranks (chunks shorter than the window and the halo included; n = 38 and 17 run
    temporal2): pooled noise sums, natively replayed thresholds, seek + halo + chunk per rank == the sequential emulated run."""
    case = fuzz.make_case(seed)
    res = fuzz.run_sharded_case(lib, case, 2 + seed % 4)
    assert res is None, (seed, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, case["cfg"], res)


def test_random_clip_stacks(fuzz, generic_lib):
    """MaxImgContainer / FastGaussianContainer kernels: aligned and unaligned buffers, chunked accumulation, 300-frame clips whose
    uint16 sums wrap like the reference's."""
    for seed in range(60):
        res = fuzz.run_stack_case(generic_lib, seed)
        assert res is None, (seed, res)


def test_random_window_readbacks(fuzz, generic_lib):
    """mdb_get_stack / mdb_get_std / mdb_get_window kernels against the checker's SlidingWindow (max, mean, sum, std, newest frame):
    windows of 1 .. 300 frames (beyond the 256-entry pointer table), rings of exactly n slots and larger, mask on the device."""
    for seed in range(40):
        res = fuzz.run_readback_case(generic_lib, seed)
        assert res is None, (seed, res)


def test_random_mfnr_clips(fuzz, tmp_path_factory):
    """Random small colour clips (2 .. 50 frames, trails in single frames, every background algorithm, random highlight / fix
    factors, random chunking) through every mfnr.cuh kernel against oracle/mfnr_oracle.py."""
    mlib = fuzz.build_mfnr_lib(str(tmp_path_factory.mktemp("emu_fuzz_mfnr")))
    for seed in range(16):
        res = fuzz.run_mfnr_case(mlib, seed)
        assert res is None, (seed, res)


@pytest.mark.parametrize("seed", [0, 2, 3])
def test_random_wide_frames(fuzz, lib, generic_lib, classic_lib, seed):
    """512 .. 1312 pixel wide frames (16 .. 41 words per row: act4's 16-byte chunks and the strip kernel for word counts that are
    not a multiple of 4), widths that are not a multiple of 32 through the generic and classic kernels."""
    case = fuzz.make_case(seed, big=True)
    gcase = fuzz.make_case(seed, any_width=True, big=True)
    res = [fuzz.run_case(lib, case, pf) for pf in (False, True)] + [fuzz.run_case(generic_lib, gcase, generic=True),
                                                                   fuzz.run_classic_case(classic_lib, gcase)]
    assert res == [None] * 4, (seed, {k: case[k] for k in ("W", "H", "n", "T", "batch")}, gcase["W"], res)


@pytest.mark.parametrize("W,H,n,T,batch,dy", [(1920, 1080, 5, 12, 6, False), (3840, 2160, 30, 34, 4, True)])
def test_baseline_frame_sizes_through_the_emulated_streaming_path(fuzz, lib, W, H, n, T, batch, dy):
    """BASELINE configs 2 and 3 at their FULL frame sizes (synthetic stream of the bench generator, a few frames beyond a full
    window) through the emulated product kernels against the checker.  The 4K case takes ~45 s: EMU_FULLSIZE=1."""
    import numpy as np
    from metdetpy_b200 import synth
    if W > 1920 and not os.environ.get("EMU_FULLSIZE"):
        pytest.skip("3840x2160, n = 30 takes ~45 s on the CPU: set EMU_FULLSIZE=1 (last run: identical)")
    fr = synth.make_stream(T, W, H, 30, speed_scale=3.0, thickness=2)
    case = dict(W=W, H=H, n=n, T=T, batch=batch, mask=np.ones((H, W), np.uint8), frames=fr, raw_frames=fr, apply_mask=False,
                cfg=dict(adaptive=True, init_value=7, sensitivity="normal", area=0.1, interval=2, hough=(10, 10, 10), dy_mask=dy))
    assert fuzz.run_case(lib, case) is None
