"""Schedule independence of the product's kernels, checked on the CPU: the block emulator (tests/emu/cuda_block_emu.h) resumes
the threads of a block in ascending order by default; EMU_SCHEDULE=1 reverses that order and EMU_SCHEDULE>=2 draws a fresh random
permutation at every round.  Every such order is a legal CUDA schedule, so the emulated PPHT (all tiers), spatial, streaming,
per-frame and time-sharded paths must reproduce the same oracle / golden results under all of them -- a kernel that does not is
missing a barrier or an atomic.  (compute-sanitizer's racecheck on the B200 is the device-side counterpart: profiles/r02_sanitizer_*.log.)"""
import os
import subprocess

import pytest

from emu_build import build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("schedule", ["1", "12345"])
def test_ppht_tiers_under_other_thread_schedules(tmp_path, schedule):
    exe = build(tmp_path, "hough_host_emu.cpp", patched=["hough.cuh"], extra_c=["oracle/ppht.c"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, EMU_SCHEDULE=schedule))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.fixture(scope="module")
def stream_lib(tmp_path_factory):
    import ctypes as C
    from emu_build import build_stream
    lib = C.CDLL(build_stream(tmp_path_factory.mktemp("stream_sched")))  # the build of tests/test_stream_emu_cpu.py when that ran first
    lib.emu_stream_path.restype = C.c_int
    return lib


@pytest.mark.parametrize("schedule", [1, 977])
def test_whole_paths_under_other_thread_schedules(stream_lib, schedule):
    """The emulated streaming / per-frame / sharded paths against the golden trajectories (the masked dy-mask case through all
    three, the dense case with tiers 2 and 3 through the streaming path) with the emulator switched to another schedule."""
    import ctypes as C
    import test_stream_emu_cpu as TS
    stream_lib.emu_set_schedule(C.c_long(schedule))
    try:
        TS.test_streaming_path_kernels_reproduce_the_reference_golden(stream_lib, "synth_384x216_n12_dyon_mask", 1000, 7)
        TS.test_streaming_path_kernels_reproduce_the_reference_golden(stream_lib, "synth_256x160_n6_fixed3_dense", 12, 5)
        TS.test_per_frame_resident_state_path_reproduces_the_reference_golden(stream_lib, "synth_384x216_n12_dyon_mask", 1000)
        TS.test_time_sharded_protocol_on_the_cpu(stream_lib, "synth_384x216_n12_dyon_mask", 3, 7)
    finally:
        stream_lib.emu_set_schedule(C.c_long(0))
