"""Debug helper: per-batch content statistics + timeline for the bench workload (not part of the product)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metdetpy_b200 import BinaryCfg, synth, _lib
from metdetpy_b200.detector import M3Detector
B = 512; W, H, n = 3840, 2160, 30
NB = int(os.environ.get("NB", "4"))
det = M3Detector(n / 30 + 1e-9, 30, np.ones((H, W), np.uint8), 10, BinaryCfg(), None, max_batch=B)
dev = torch.device("cuda", 0)
xs = [synth.make_stream_device(B, W, H, 30, dev, t0=s * B) for s in range(NB)]
torch.cuda.synchronize()
for s in range(3):
    det.submit(xs[s % NB].data_ptr(), B, True); det.collect()
det._eng.set_option("timeline", 1)
lib = _lib.load(); lib.mdb_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
out = np.zeros(9, np.float32)
det.submit(xs[3 % NB].data_ptr(), B, True)
names = ["front0", "thr_done", "temporal0", "act_done", "dst_done", "hough1_done", "hough_done", "copied", "dst0"]
for s in range(3, 3 + 2 * NB):
    det.submit(xs[(s + 1) % NB].data_ptr(), B, True)
    res = det.collect()
    li = det.last_infos
    lib.mdb_debug_timeline(det._eng.handle, out.ctypes.data)
    non = li["n_on"]; ln = li["lines_num"]
    print(f"batch {s % NB}: n_on max={non.max()} >2048:{int((non > 2048).sum())} >4096:{int((non > 4096).sum())} "
          f"sum={int(non.sum())} lines max={ln.max()} thr={np.unique(li['bi_threshold'])} | "
          + "  ".join(f"{k}={v:.2f}" for k, v in zip(names, out)))
det.collect()
