"""Dense-mask regime: Hough tiers 2/3 against cv2 on the same masks (exactness + time).  4K noise masks with a fixed
low threshold; compares raw segments of a few frames with cv2.HoughLinesP run on the device's own dst masks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import bench
from metdetpy_b200.detector import M3Detector
W, H, n = 3840, 2160, 30
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
st = bench.Stream(64, W, H, 30.0, dev, 0, distinct=2, quiet=n)
for thr in (6, 5):
    det = M3Detector(n / 30 + 1e-9, 30.0, np.ones((H, W), np.uint8), 10, bench.make_cfg(dy=True, adaptive=False, init_value=thr),
                     None, max_batch=64)
    det._eng.set_option("hough_profile", 1)
    det.detect_many((st.ptr(0), 64), on_device=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, dst = det.detect_many((st.ptr(64), T), on_device=True, return_dst=True)
    dt = time.perf_counter() - t0
    info = det.last_infos
    print(f"thr {thr}: {T} frames in {dt * 1e3:.1f} ms = {T / dt:.1f} frames/s; on-pixels mean {info['n_on'].mean():.0f}, "
          f"raw segments mean {info['lines_num'].mean():.1f}", flush=True)
    import ctypes as C
    from metdetpy_b200 import _lib
    prof = np.zeros((64, 10), np.int64)
    assert _lib.load().mdb_debug_hough_profile(det._eng.handle, prof.ctypes.data, 64) == 0
    m = prof[:T].mean(0) / 1.965e6
    print(f"  per-frame PPHT phases (ms @1.965 GHz): setup {m[1]:.1f} vote {m[2]:.1f} walk {m[3]:.1f} second pass/unvote {m[4]:.1f} "
          f"staging {m[5]:.1f} total {m[8]:.1f}; points voted {prof[:T, 6].mean():.0f}, triggers {prof[:T, 7].mean():.0f}, "
          f"isolated {prof[:T, 9].mean():.0f}", flush=True)
    bad = 0
    tc = 0.0
    for i in range(min(T, 6)):
        gap = float(info["gap"][i])
        t1 = time.perf_counter()
        ref = cv2.HoughLinesP(dst[i], 1, np.pi / 180, 10, minLineLength=10, maxLineGap=gap)
        tc += time.perf_counter() - t1
        ref = np.zeros((0, 4), np.int32) if ref is None else ref.reshape(-1, 4)
        got = det.last_raw[i].reshape(-1, 4) if info["lines_num"][i] <= 500 else None
        if info["lines_num"][i] != len(ref) or (got is not None and not np.array_equal(got, ref)):
            bad += 1
            print("  MISMATCH frame", i, info["lines_num"][i], len(ref))
    print(f"  cv2.HoughLinesP alone: {tc / min(T, 6) * 1e3:.1f} ms per frame; mismatching frames: {bad}", flush=True)
    det.close()
