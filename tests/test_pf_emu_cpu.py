"""The per-frame resident-state kernels (csrc/perframe_kernel.cuh) on the HOST: tests/emu/pf_host_emu.cpp includes the
kernel source with the CUDA built-ins emulated and drives it the way csrc/metdet.cu does (staging buffer, ring slot written
by the update kernel in two halves, suffix planes at block ends, rebuild from the ring after a jump), against a brute-force
statement of SlidingWindow.update + the predicate max*L - sum > thr*L (MetLib/utils.py:269-307, Detector.py:327-332):
windows 2..30, ring of exactly n slots and larger, masked and unmasked, three kinds of content.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_per_frame_kernels_against_brute_force(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include"))
                     if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "pf_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", cuda_inc, os.path.join(REPO, "tests", "emu", "pf_host_emu.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
