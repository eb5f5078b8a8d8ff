"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): streaming + generic
kernels, dy-mask, masked loads, Hough tiers 1a and 3, per-frame API, max stack."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth, stacker
from metdetpy_b200.detector import M3Detector
W, H, FPS, n, T = 256, 96, 30, 5, 24
rng_m = np.random.default_rng(4)
frames = synth.make_stream(T, W, H, FPS, speed_scale=3.0, thickness=2)
mask = np.ones((H, W), np.uint8); mask[80:, :] = 0
cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 1), HoughLineCfg(8, 8, 6), DynamicCfg(True, 5))
for mode in ("stream", "temporal_v3", "stream_dense_dst", "stream_strip_act", "generic"):
    det = M3Detector(n / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=8, apply_mask=True)
    det._eng.set_option("stream_kernel", int(mode != "generic"))
    det._eng.set_option("force_dense", int(mode == "stream_dense_dst"))
    det._eng.set_option("force_strip", int(mode == "stream_strip_act"))
    det._eng.set_option("temporal_version", 3 if mode == "temporal_v3" else 2)
    tot = 0
    for s in range(0, T, 8):
        res, dst = det.detect_many(frames[s:s + 8], return_dst=True)
        tot += sum(len(r[0]) for r in res)
    print(mode, "lines", tot)
    det.close()
# sub-blocked temporal kernel (window 6 = 3 blocks of 2)
det = M3Detector(6 / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=8, apply_mask=True)
det._eng.set_option("temporal_kdiv", 3)
print("subblocks lines", sum(len(r[0]) for s in range(0, T, 8) for r in det.detect_many(frames[s:s + 8])))
det.close()
det = M3Detector(n / FPS + 1e-9, FPS, mask, 10, cfg, None)
for t in range(8):
    det.update(frames[t] * mask); det.detect()
_ = det.dst, det.stack.max
# temporal3 shapes: paged shared-memory ring (n = 30), long window (n = 60), two far-apart objects (two rho intervals)
for nn in (30, 60):
    fr = synth.make_stream(80, W, H, FPS, speed_scale=3.0, thickness=2)
    d3 = M3Detector(nn / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=40, apply_mask=True)
    print("temporal3 n", nn, "lines", sum(len(r[0]) for s in range(0, 80, 40) for r in d3.detect_many(fr[s:s + 40])),
          "generation", int(d3._eng.info("temporal_generation")))
    d3.close()
# per-frame API on the resident-state path across two block ends, a batched call in between (state rebuild), read-backs
dp = M3Detector(n / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=4, apply_mask=True)
for t in range(12):
    dp.update(frames[t]); dp.detect()
dp.detect_many(frames[12:16])
for t in range(16, 22):
    dp.update(frames[t]); dp.detect()
print("per-frame generation", int(dp._eng.info("temporal_generation")), int(dp.stack.max.sum()), int(dp.stack.sliding_window.sum()))
dp.close()
# MFNR mix stacker, background algorithms mean / sigma clipping / median, odd and aligned sizes
for shp in ((40, 52, 3), (37, 45, 3)):
    clipf = rng_m.integers(0, 256, (9,) + shp, dtype=np.uint8)
    for algo in ("mean", "sigma-clipping", "median"):
        bx = stacker.MfnrMixContainer(keep_frames=algo != "mean", chunk=4)
        for f in clipf:
            bx.append(f)
        print("mfnr", shp, algo, int(bx.export(0.9, 31, algo, 1.5).sum()))
        bx.close()
# dense frame -> tier 3
rng = np.random.default_rng(0)
dense = rng.integers(0, 60, (6, H, W)).astype(np.uint8)
cfg2 = BinaryCfg(BinaryCoreCfg(False, 10, "normal", 0.2, 1), HoughLineCfg(10, 10, 10), DynamicCfg(False, 5))
d2 = M3Detector(3 / FPS + 1e-9, FPS, np.ones((H, W), np.uint8), 10, cfg2, None, max_batch=6)
d2.detect_many(dense)
print("dense n_on", [i["n_on"] for i in d2.last_infos])
print("maxstack", stacker.merge_max(frames[:5]).sum())
# streaming Gaussian stack: aligned (16-byte) and odd-sized frames, accumulation across chunks
for shp in ((32, 48, 3), (7, 11, 3)):
    box = stacker.FastGaussianContainer(chunk=3)
    for f in rng.integers(0, 256, (8,) + shp, dtype=np.uint8):
        box.append(f)
    print("gauss", shp, int(box.container.sum_mu.sum()))
# loader preprocessing (resize -> gray -> mask -> exposure merge), ClassicDetector (aligned and odd widths)
from metdetpy_b200.imgproc import Transform
from metdetpy_b200.detector import ClassicDetector
bgr = rng.integers(0, 256, (5, 50, 70, 3), dtype=np.uint8)
tr = Transform(); tr.opencv_resize([33, 21]); tr.opencv_BGR2GRAY(); tr.mask_with(np.ones((21, 33), np.uint8))
print("preproc", tr.exec_transform_many(bgr, 2).sum()); tr.close()
for Wc in (256, 203):
    fr = synth.make_stream(12, Wc, 96, FPS, speed_scale=3.0, thickness=2)
    cd = ClassicDetector(1.0, FPS, np.ones((96, Wc), np.uint8), 10, cfg, None, max_batch=5)
    n = 0
    for s in range(0, 12, 5):
        n += sum(len(r[0]) for r in cd.detect_many(fr[s:s + 5]))
    cd.update(fr[0]); cd.detect(); _ = cd.dst
    print("classic", Wc, "lines", n); cd.close()
