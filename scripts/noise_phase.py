"""Times phase A of the time-sharded path (noise sums of one rank's chunk, mdb_noise_sums_dev) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from metdetpy_b200 import sharding as S
from metdetpy_b200.detector import M3Detector
W, H, n, B = 3840, 2160, 30, 512
dev = torch.device("cuda", 0)
det = M3Detector(n / 30 + 1e-9, 30.0, np.ones((H, W), np.uint8), 10, bench.make_cfg(), None, max_batch=B)
st = bench.Stream(B, W, H, 30.0, dev, 2 * n - 2, distinct=2, quiet=n)
C = 240 * B
segs = [S.Segment(st.ptr(t), B, t, history=2 * n - 2) for t in range(C, 2 * C, B)]
for r in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = S.chunk_noise_samples(det, segs, n, 2, C, 2 * C, W * H)
    print(f"chunk_noise_samples: {(time.perf_counter() - t0) * 1e3:.2f} ms for {len(out)} samples")
