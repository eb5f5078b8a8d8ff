"""Streaming Gaussian stack on the device (first half): GB/s of gauss_stack_kernel on 4K BGR frames
next to numpy (the reference's adds) on the host.  Prints one JSON object (a side measurement)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metdetpy_b200 import _lib
lib = _lib.load()
T, H, W, C = 64, 2160, 3840, 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
frames = torch.randint(0, 256, (T, H, W, C), dtype=torch.uint8, device=dev, generator=g)
s = torch.empty((H, W, C), dtype=torch.int16, device=dev); q = torch.empty((H, W, C), dtype=torch.int32, device=dev)
fb = H * W * C
ts = []
for _ in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    assert lib.mdb_gauss_stack(frames.data_ptr(), T, fb, s.data_ptr(), q.data_ptr(), 1, 1, 0, 0) == 0
    ts.append(time.perf_counter() - t0)
dt = float(np.median(ts[1:]))
host = frames[:6].cpu().numpy()
t0 = time.perf_counter()
acc = host[0].astype(np.uint16); sq = np.square(acc, dtype=np.uint32)
for f in host[1:]:
    x = f.astype(np.uint16); acc = acc + x; sq = sq + np.square(x, dtype=np.uint32)
cpu = (len(host)) / (time.perf_counter() - t0)
peak = 6650.0
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
bytes_alg = T * fb + 6 * fb
print(json.dumps({"workload": f"{T} device-resident {W}x{H} BGR frames -> sum (u16) + sum of squares (u32)",
                  "call_ms": dt * 1e3, "frames_per_s": T / dt, "GBps": bytes_alg / dt / 1e9, "peak_GBps": peak,
                  "frac": bytes_alg / dt / 1e9 / peak, "cpu_numpy_frames_per_s": cpu}))
