"""The spatial half of the mask chain (csrc/spatial_kernel.cuh: act4_kernel and the warp-strip act_kernel, dst_sparse_kernel,
dst_dense_kernel, the persistent mask buffer with its shadow bits and word lists) run on the CPU by the thread-block emulator
with the product's launch geometry, over several batches that reuse the same buffers, against a per-pixel statement of
MetLib/Detector.py:329-335 (median, threshold, close) and :234-242 (dynamic mask: not on in ALL of the last L act frames,
eroded; dst = act * m): mask bytes, on-pixel counts and lists, shadow bits.  No GPU needed."""
import subprocess

from emu_build import build


def test_spatial_kernels_against_per_pixel_reference(tmp_path):
    exe = build(tmp_path, "spatial_host_emu.cpp", patched=["spatial_kernel.cuh"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
