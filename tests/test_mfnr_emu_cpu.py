"""The frame-pass kernels of the MFNR stacker (csrc/mfnr.cuh: accumulation, sigma clipping, medians) on the HOST:
tests/emu/mfnr_host_emu.cpp includes the kernel source with the CUDA built-ins emulated and checks every element against
brute-force statements of the reference's containers (MetLib/stacker.py:43-59, uint16 / uint32 wrap-around incl. a 300-frame
clip that wraps), single_sigma_clipping (:94-115) and np.median / median_of_medians (:62-78).  No GPU needed."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mfnr_frame_pass_kernels_against_brute_force(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include"))
                     if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "mfnr_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", cuda_inc,
                           os.path.join(REPO, "tests", "emu", "mfnr_host_emu.cpp"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
