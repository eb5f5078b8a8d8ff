"""act4_kernel (csrc/spatial_kernel.cuh) on the HOST: tests/emu/act4_host_emu.cpp includes the kernel source with the CUDA
built-ins emulated and compares the act bit-frame with a per-pixel statement of cv2.medianBlur(., 3) (border replicated) ->
threshold -> cv2.morphologyEx(MORPH_CLOSE, 3x3 rect) (outside pixels ignored), MetLib/Detector.py:329-335, plus the list
of non-zero words the kernel emits: widths 128..384, heights 1..67, band heights 8 and 64, empty to full masks.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_act4_kernel_against_per_pixel_median_and_close(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include"))
                     if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "act4_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", cuda_inc, os.path.join(REPO, "tests", "emu", "act4_host_emu.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
