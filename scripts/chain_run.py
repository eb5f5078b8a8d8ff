"""Runs a few batches of one BASELINE configuration with ONE batch in flight (the mask chain alone on the GPU), for
ncu captures of the chain kernels.  usage: python scripts/chain_run.py W H n B [dy=1] [mask=0] [batches=4]
Prints the library's own CUDA-event times of the temporal pass and of the whole chain."""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
import bench
from metdetpy_b200.detector import M3Detector

a = sys.argv[1:]
W, H, n, B = (int(v) for v in a[:4])
dy = bool(int(a[4])) if len(a) > 4 else True
mk = bool(int(a[5])) if len(a) > 5 else False
nb = int(a[6]) if len(a) > 6 else 4
fps = 60.0 if n == 60 else 30.0
dev = torch.device("cuda", 0)
mask = np.ones((H, W), np.uint8)
if mk:
    mask, note = bench.load_bench_mask(W, H)
    print("mask:", note)
det = M3Detector(n / fps + 1e-9, fps, mask, 10, bench.make_cfg(dy=dy), None, max_batch=B, apply_mask=mk)
st = bench.Stream(B, W, H, fps, dev, 0, distinct=2)
for s in range(nb):
    det.submit(st.ptr(s * B), B, True)
    det.collect(want_lines=False)
    print(f"batch {s}: temporal {det._eng.info('temporal_ms'):.3f} ms, chain {det._eng.fused_time()[0]:.3f} ms, "
          f"generation {int(det._eng.info('temporal_generation'))}", flush=True)
det.close()
