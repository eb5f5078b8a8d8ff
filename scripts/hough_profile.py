"""Debug helper: per-frame PPHT phase cycle counters for one 4K batch (not part of the product)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metdetpy_b200 import BinaryCfg, synth, _lib
from metdetpy_b200.detector import M3Detector
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H, n = 3840, 2160, 30
det = M3Detector(n / 30 + 1e-9, 30, np.ones((H, W), np.uint8), 10, BinaryCfg(), None, max_batch=B)
det._eng.set_option("hough_profile", 1)
dev = torch.device("cuda", 0)
for s in range(3):
    x = synth.make_stream_device(B, W, H, 30, dev, t0=s * B)
    torch.cuda.synchronize()
    det.submit(x.data_ptr(), B, True); det.collect()
out = np.zeros((B, 10), np.int64)
lib = _lib.load(); lib.mdb_debug_hough_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert lib.mdb_debug_hough_profile(det._eng.handle, out.ctypes.data, B) == 0
np.set_printoptions(linewidth=200)
print("N setup vote walk unvote [tier1: loads+sort inside setup | tier2: reset | tier3: staging] n_vote n_line total lines  (cycles)")
o = out[np.argsort(-out[:, 8])]
print(o[:12]); print("mean", out.mean(0).astype(int)); print("sum-of-total ms @1.9GHz", out[:, 8].sum() / 1.9e6, "max", out[:, 8].max() / 1.9e6)
