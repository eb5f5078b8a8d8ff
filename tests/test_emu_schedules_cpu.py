"""Schedule independence of the product's kernels, checked on the CPU: the block emulator (tests/emu/cuda_block_emu.h) resumes
the threads of a block in ascending order by default; EMU_SCHEDULE=1 reverses that order and EMU_SCHEDULE>=2 draws a fresh random
permutation at every round.  Every such order is a legal CUDA schedule, so the emulated PPHT (all tiers), spatial, streaming,
per-frame and time-sharded paths must reproduce the same oracle / golden results under all of them -- a kernel that does not is
missing a barrier or an atomic.  (compute-sanitizer's racecheck on the B200 is the device-side counterpart: profiles/r02_sanitizer_*.log.)"""
import os
import subprocess
import sys

import pytest

from emu_build import build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("schedule", ["1", "12345"])
def test_ppht_tiers_under_other_thread_schedules(tmp_path, schedule):
    exe = build(tmp_path, "hough_host_emu.cpp", patched=["hough.cuh"], extra_c=["oracle/ppht.c"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, EMU_SCHEDULE=schedule))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("schedule", ["1", "977"])
def test_whole_paths_under_other_thread_schedules(schedule):
    """The emulated streaming / per-frame / sharded paths against the golden trajectories (the masked dy-mask case through all three, the dense case with
    tiers 2 and 3 through the streaming path) in a child process whose emulator uses another schedule."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REPO, "tests", "test_stream_emu_cpu.py"), "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "synth_384x216 or (dense and streaming_path)"],
                       capture_output=True, text=True, timeout=1800, cwd=REPO, env=dict(os.environ, EMU_SCHEDULE=schedule))
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
