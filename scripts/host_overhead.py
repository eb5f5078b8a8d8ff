"""Debug helper: where does the host time of one pipelined step go? (not part of the product)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from metdetpy_b200 import BinaryCfg, synth
from metdetpy_b200.detector import M3Detector, _ptr
from metdetpy_b200._lib import check
B = 512; W, H, n = 3840, 2160, 30
det = M3Detector(n / 30 + 1e-9, 30, np.ones((H, W), np.uint8), 10, BinaryCfg(), None, max_batch=B)
dev = torch.device("cuda", 0)
xs = [synth.make_stream_device(B, W, H, 30, dev, t0=s * B) for s in range(3)]
torch.cuda.synchronize()
eng = det._eng
for s in range(3):
    det.submit(xs[s % 3].data_ptr(), B, True); det.collect()
ts = dict(submit=0.0, c_collect=0.0, unpack=0.0)
t_all = time.perf_counter()
det.submit(xs[0].data_ptr(), B, True)
N = 10
for s in range(N):
    t0 = time.perf_counter(); det.submit(xs[(s + 1) % 3].data_ptr(), B, True); t1 = time.perf_counter()
    T = det._pending.pop(0)
    check(eng.lib.mdb_collect_batch(eng.handle, C.byref(eng.infos), _ptr(eng.lines), _ptr(eng.prob), _ptr(eng.raw), None, 0), "collect")
    t2 = time.perf_counter(); res = det._unpack_all(T); t3 = time.perf_counter()
    ts["submit"] += t1 - t0; ts["c_collect"] += t2 - t1; ts["unpack"] += t3 - t2
det.collect()
print("per step ms:", {k: round(v / N * 1e3, 3) for k, v in ts.items()}, "wall/step", round((time.perf_counter() - t_all) / (N + 1) * 1e3, 3))
