"""Debug helper: event timeline of update(); detect() on one 4K frame (per-frame O(1) path), device-resident input."""
import os, sys, ctypes as C, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from metdetpy_b200 import _lib
from metdetpy_b200._lib import check
from metdetpy_b200.detector import M3Detector
W, H, n, F = 3840, 2160, 30, 90
dev = torch.device("cuda", 0)
st = bench.Stream(F, W, H, 30.0, dev, 0, distinct=1)
det = M3Detector(n / 30.0 + 1e-9, 30.0, np.ones((H, W), np.uint8), 10, bench.make_cfg(), None)
lib = _lib.load()
def step(t):
    check(lib.mdb_update(det._eng.handle, st.ptr(t), 1), "update"); det._timer += 1
    return det.detect()
for t in range(50):
    step(t)
if len(sys.argv) > 1:  # plain loop for an ncu launch list
    for t in range(50, 60):
        step(t)
    sys.exit(0)
det._eng.set_option("timeline", 1)
out = np.zeros(9, np.float32)
names = ["front0", "thr_done", "ev_f0", "ev_f1", "hough0", "tier1_done", "tiers_done", "copied", "ev_d0"]
t = 50
for label, opts in (("default", {}), ("single_dense=0", {"single_dense": 0}), ("sp_rows_single=64", {"single_dense": 1, "sp_rows_single": 64})):
    for k, v in opts.items():
        det._eng.set_option(k, v)
    print(label)
    for _ in range(8):
        a = time.perf_counter(); step(t); b = time.perf_counter()
        lib.mdb_debug_timeline(det._eng.handle, out.ctypes.data)
        base = out[0]
        print(f"frame {t}: host {1e6 * (b - a):6.1f} us | " + "  ".join(f"{k}={1e3 * (v - base):6.1f}" for k, v in zip(names, out)) +
              f" | n_on {det._eng.infos[0].n_on}")
        t += 1
