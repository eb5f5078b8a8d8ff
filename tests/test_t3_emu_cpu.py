"""temporal3_kernel's per-thread code on the HOST: tests/emu/t3_host_emu.cpp includes csrc/temporal3_kernel.cuh with its
intrinsics emulated (T3_HOST_EMU) and runs every thread of a small frame serially against a brute-force statement of the
predicate max(window)*L - sum(window) > thr*L (MetLib/utils.py:269-307, MetLib/Detector.py:327-332): register ring, shared
page, sub-blocked van Herk, window sums, thresholds per frame, the mask folded into the bit-gather weights, warm-up (t0 = 0),
ragged batch lengths -- for the default shapes of n = 5 / 30 / 60 and a dozen others.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_temporal3_thread_code_against_brute_force(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include"))
                     if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "t3_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", cuda_inc, os.path.join(REPO, "tests", "emu", "t3_host_emu.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") >= 15
