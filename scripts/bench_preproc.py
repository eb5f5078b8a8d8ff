"""Loader preprocessing on the device: frames/s and GB/s of preproc_kernel on 4K BGR
frames -> 960x540 gray (the reference's default runtime size), next to cv2 on the host cores.
Prints one JSON object (a side measurement; bench.py's line is the headline)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from metdetpy_b200.imgproc import Transform

W0, H0, W, H, T, EXP = 3840, 2160, 960, 540, 64, 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
frames = torch.randint(0, 256, (T, H0, W0, 3), dtype=torch.uint8, device=dev, generator=g)
mask = np.ones((H, W), np.uint8)
tr = Transform()
tr.opencv_resize([W, H]); tr.opencv_BGR2GRAY(); tr.mask_with(mask)
ms = []
for it in range(8):
    tr.exec_transform_many((frames.data_ptr(), T), EXP, on_device=True, keep_on_device=True, shape=(H0, W0, 3))
    ms.append(tr.last_kernel_ms())
ms = float(np.median(ms[2:]))
src_bytes = T * H0 * W0 * 3
out_bytes = (T // EXP) * H * W
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
host = frames[:8].cpu().numpy()
t0 = time.perf_counter()
for f in host:
    _ = cv2.cvtColor(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) * mask
cpu_fps = len(host) / (time.perf_counter() - t0)
# 4x downscale touches 2 of every 4 source rows: bytes a perfect kernel must read = rows used * row bytes
from metdetpy_b200 import _lib as _L  # the library's own tap table (mdb_preproc_axis_taps): source rows the vertical taps touch
_taps = [np.zeros(H, np.int32) for _ in range(4)]
_L.check(_L.load().mdb_preproc_axis_taps(H, H0, 0, *[a.ctypes.data for a in _taps]), "axis taps")
rows_used = len(set(np.concatenate(_taps[:2]).tolist()))
need = T * rows_used * W0 * 3 + out_bytes
print(json.dumps({"workload": f"{T} frames {W0}x{H0} BGR -> {W}x{H} gray, mask, exp_frame={EXP}, device-resident",
                  "kernel_ms": ms, "frames_per_s": T / (ms * 1e-3),
                  "source_GBps": src_bytes / (ms * 1e-3) / 1e9, "needed_bytes_GBps": need / (ms * 1e-3) / 1e9,
                  "peak_GBps": peak, "frac_of_peak_needed_bytes": need / (ms * 1e-3) / 1e9 / peak,
                  "cpu_cv2_frames_per_s": cpu_fps, "cpu_threads": cv2.getNumThreads()}))
