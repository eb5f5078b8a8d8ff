"""The exact-order PPHT kernels of the product (csrc/hough.cuh: visiting order, tiers 1a / 1b in shared memory, tier 2 and
the dense tier 3 in global memory) run on the CPU by the thread-block emulator (tests/emu/cuda_block_emu.h) and compared,
segment for segment, with oracle/ppht.c -- the restatement of cv2.HoughLinesP (MetLib/Detector.py:347-352) that
tests/test_oracle_golden.py pins on cv2 itself.  Twelve masks from 97x61 to 3840x2160: thin and thick lines, noise, two
far-apart objects (two rho intervals per angle), > 4096 points (tier 2), > 16384 points (tier 3).  No GPU needed."""
import subprocess

from emu_build import build


def test_ppht_kernels_equal_the_oracle_on_the_cpu(tmp_path):
    exe = build(tmp_path, "hough_host_emu.cpp", patched=["hough.cuh"], extra_c=["oracle/ppht.c"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    tiers = {line.split("tier ")[1].split(":")[0] for line in r.stdout.splitlines() if "tier " in line}
    assert tiers == {"1a", "1b", "2", "3"}, tiers
