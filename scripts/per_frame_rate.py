"""update(frame); detect() per frame at 4K from pinned host memory (MetDetPy.py:197-198): frames/s of the O(1) path
(resident window state) and of the window-re-reading path, plus the device-resident rate (no PCIe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from metdetpy_b200.detector import M3Detector
from metdetpy_b200._lib import check

W, H, n, F = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 30, 160
dev = torch.device("cuda", 0)
st = bench.Stream(F, W, H, 30.0, dev, 0, distinct=1)
host = torch.empty((F, H, W), dtype=torch.uint8).pin_memory()
host.copy_(st.view(0, F))
frames = host.numpy()
for fast in (1, 0):
    det = M3Detector(n / 30.0 + 1e-9, 30.0, np.ones((H, W), np.uint8), 10, bench.make_cfg(), None)
    det._eng.set_option("per_frame_fast", fast)
    for mode in ("host", "device"):
        det.reset()
        for t in range(40):
            det.update(frames[t]); det.detect()
        torch.cuda.synchronize()
        tu = td = 0.0
        t0 = time.perf_counter()
        for t in range(40, F):
            a = time.perf_counter()
            if mode == "host":
                det.update(frames[t])
            else:
                check(det._eng.lib.mdb_update(det._eng.handle, st.ptr(t), 1), "update"); det._timer += 1
            b = time.perf_counter()
            det.detect()
            c = time.perf_counter()
            tu += b - a; td += c - b
        dt = time.perf_counter() - t0
        k = F - 40
        print(f"per_frame_fast={fast} {mode:6s}: {k / dt:8.1f} frames/s  ({1e6 * dt / k:6.1f} us/frame: update() {1e6 * tu / k:6.1f} us, "
              f"detect() {1e6 * td / k:6.1f} us; generation {int(det._eng.info('temporal_generation'))})", flush=True)
    det.close()
