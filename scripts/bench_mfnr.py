"""MFNR mix stacker on a device-resident 4K colour clip: wall time of append x3 + finish (sigma clipping), for ncu launch lists."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metdetpy_b200 import _lib, stacker
from metdetpy_b200._lib import check
dev = torch.device("cuda", 0)
Tm, Hm, Wm = 48, 2160, 3840
g = torch.Generator(device=dev); g.manual_seed(3)
clip = (torch.randn((Tm, Hm, Wm, 3), device=dev, generator=g) * 5.0 + 50.0).clamp_(0, 255).to(torch.uint8)
clip[:, 1000:1003, 500:2500] = 240
torch.cuda.synchronize()
lib = _lib.load()
for algo in (1, 0):
    for rep in range(3):
        h = C.c_void_p()
        check(lib.mdb_mfnr_create(Hm, Wm, 3, algo, 0, C.byref(h)), "mfnr")
        prm = _lib.MfnrParams()
        prm.highlight_preserve, prm.blur_ksize, prm.blur_sigma, prm.bg_algorithm = 0.9, 31, 3.0, algo
        prm.sigma_high = prm.sigma_low = 3.0
        prm.bg_fix_factor, prm.gumbel_mean = 1.5, float(stacker.get_gumbel_mean(Tm))
        outb = torch.empty((Hm, Wm, 3), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        check(lib.mdb_mfnr_reserve(h, Tm), "mfnr")
        t1 = time.perf_counter()
        for s0 in range(0, Tm, 16):
            check(lib.mdb_mfnr_append(h, clip[s0:s0 + 16].data_ptr(), 16, 1), "mfnr")
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        check(lib.mdb_mfnr_finish(h, C.byref(prm), outb.data_ptr(), 1, None), "mfnr")
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        lib.mdb_mfnr_destroy(h)
        print(f"algo {algo} rep {rep}: reserve {1e3 * (t1 - t0):.2f} ms, append {1e3 * (t2 - t1):.2f} ms, finish {1e3 * (t3 - t2):.2f} ms, total {1e3 * (t3 - t0):.2f} ms")
