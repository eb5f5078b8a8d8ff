#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the LIVE, UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the GPU box gets the committed .npz
files.  Usage:   python tests/golden/make_golden.py [case ...]

What is recorded, per detector case (real `MetLib.Detector.M3Detector`, reference file
MetLib/Detector.py:302-392): the input frames, the mask, and for every frame the trajectory
`bi_threshold`, `bi_threshold_float`, `snr`, the binary mask `dst` (bit-packed), `dst_sum`,
the raw HoughLinesP segments (`linesp_ext`), `lines_num`, the NMS'd lines and `cls_pred`.
Plus: `lineset_nms` (MetLib/utils.py:780-839) on random segment sets, a `SlidingWindow`
trace (MetLib/utils.py:225-321), `SNR_SW.select_subarea` ROIs (MetLib/Detector.py:93-122),
and raw `cv2.HoughLinesP` outputs on assorted masks (the third-party call at Detector.py:347).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

import cv2  # noqa: E402
import numpy as np  # noqa: E402

from MetLib.Detector import M3Detector, SNR_SW  # noqa: E402
from MetLib.metlog import BaseMetLog  # noqa: E402
from MetLib.metstruct import (BinaryCfg, BinaryCoreCfg, DynamicCfg,  # noqa: E402
                              HoughLineCfg)
from MetLib.utils import EMA, SlidingWindow, lineset_nms  # noqa: E402

from metdetpy_b200 import synth  # noqa: E402


def ragged(list_of_arrays, width, dtype):
    offs = np.zeros(len(list_of_arrays) + 1, np.int64)
    for i, a in enumerate(list_of_arrays):
        offs[i + 1] = offs[i] + len(a)
    flat = np.zeros((offs[-1], width), dtype)
    for i, a in enumerate(list_of_arrays):
        if len(a):
            flat[offs[i]:offs[i + 1]] = np.asarray(a).reshape(-1, width)
    return flat, offs


def run_detector_case(name, frames, mask, n, fps, cfg_tuple, hough=(10, 10, 10), dy=True):
    adaptive, init_value, sens, area, interval = cfg_tuple
    cfg = BinaryCfg(BinaryCoreCfg(adaptive, init_value, sens, area, interval),
                    HoughLineCfg(*hough), DynamicCfg(dy, 5))
    det = M3Detector(window_sec=n / fps + 1e-9, fps=fps, mask=mask, num_cls=10, cfg=cfg,
                     logger=BaseMetLog())
    assert det.stack_maxsize == n, (det.stack_maxsize, n)
    T, H, W = frames.shape
    thr, thrf, snr, dsts, dsum = [], [], [], [], []
    raw, nraw, nms, cls = [], [], [], []
    for t in range(T):
        det.update(frames[t])
        lines, cp = det.detect()
        thr.append(int(det.bi_threshold))
        thrf.append(float(det.bi_threshold_float))
        snr.append(float(det.stack.snr))
        dsts.append(np.packbits(det.dst > 0, axis=None))
        assert set(np.unique(det.dst)) <= {0, 255}
        dsum.append(float(det.dst_sum))
        raw.append(np.asarray(det.linesp_ext).reshape(-1, 4))
        nraw.append(int(det.lines_num))
        nms.append(np.asarray(lines).reshape(-1, 4))
        cls.append(np.asarray(cp, np.float64).reshape(-1, 10))
    raw_f, raw_o = ragged(raw, 4, np.int32)
    nms_f, nms_o = ragged(nms, 4, np.int32)
    cls_f, _ = ragged(cls, 10, np.float64)
    out = dict(frames=frames, mask=mask, n=n, fps=float(fps),
               cfg_adaptive=adaptive, cfg_init_value=init_value, cfg_sensitivity=sens,
               cfg_area=float(area), cfg_interval=interval, hough=np.array(hough, np.int64),
               dy_mask=dy, std_roi=np.array(det.stack.std_roi, np.int64),
               mask_area=int(det.mask_area),
               bi_threshold=np.array(thr, np.int64), bi_threshold_float=np.array(thrf),
               snr=np.array(snr), dst_bits=np.stack(dsts), dst_sum=np.array(dsum),
               raw_lines=raw_f, raw_offs=raw_o, lines_num=np.array(nraw, np.int64),
               nms_lines=nms_f, nms_offs=nms_o, cls_pred=cls_f)
    path = os.path.join(HERE, f"det_{name}.npz")
    np.savez_compressed(path, **out)
    nz = [int(np.unpackbits(d).sum()) for d in dsts]
    print(f"{name}: T={T} {W}x{H} n={n} thr {thr[0]}..{thr[-1]} max_nnz={max(nz)} "
          f"frames_with_lines={sum(1 for a in nms if len(a))} max_raw={max(nraw)} "
          f"-> {os.path.getsize(path) / 1e6:.2f} MB")


NORMAL = (True, 7, "normal", 0.1, 2)


def case_synth_small():
    W, H, FPS, n = 320, 240, 30, 5
    fr = synth.make_stream(75, W, H, FPS, speed_scale=3.0, thickness=2)
    run_detector_case("synth_320x240_n5_dyoff", fr, np.ones((H, W), np.uint8), n, FPS, NORMAL,
                      dy=False)


def case_synth_dy_mask():
    W, H, FPS, n = 384, 216, 24, 12
    fr = synth.make_stream(90, W, H, FPS, speed_scale=2.5, thickness=2)
    mask = np.ones((H, W), np.uint8)
    mask[150:, :] = 0          # "ground" at the bottom
    mask[60:130, 120:200] = 0  # a blocked patch inside the centre ROI -> ROI slides up
    fr = fr * mask[None]       # loader's mask_with (imgproc.py:96-101)
    run_detector_case("synth_384x216_n12_dyon_mask", fr, mask, n, FPS, NORMAL, dy=True)


def case_odd_size():
    W, H, FPS, n = 203, 157, 30, 3
    fr = synth.make_stream(40, W, H, FPS, speed_scale=4.0, thickness=2)
    run_detector_case("synth_203x157_n3_high", fr, np.ones((H, W), np.uint8), n, FPS,
                      (True, 7, "high", 0.1, 1), hough=(8, 8, 6), dy=True)


def case_fixed_thr_dense():
    """Non-adaptive low threshold on noisy frames: dense masks, > 500 raw lines on some frames
    (NUM_LINES_TOOMUCH guard, Detector.py:358-360), dynamic gap at its floor."""
    W, H, FPS, n = 256, 160, 30, 6
    fr = synth.make_stream(30, W, H, FPS, speed_scale=3.0, thickness=1, sigma=3.0)
    run_detector_case("synth_256x160_n6_fixed3_dense", fr, np.ones((H, W), np.uint8), n, FPS,
                      (False, 3, "normal", 0.1, 2), dy=True)


def case_low_sens():
    W, H, FPS, n = 300, 200, 25, 7
    fr = synth.make_stream(60, W, H, FPS, speed_scale=3.0, thickness=2)
    run_detector_case("synth_300x200_n7_low", fr, np.ones((H, W), np.uint8), n, FPS,
                      (True, 7, "low", 0.2, 1), dy=True)


def case_real_clip():
    """Crop of the bundled clip (test/20220413Red.mp4) around its meteor, decoded by cv2/FFmpeg,
    resize (960,540) INTER_LINEAR -> BGR2GRAY (the loader's Transform, videoloader.py:300-308)."""
    cap = cv2.VideoCapture("/root/reference/test/20220413Red.mp4")
    x0, y0, w, h = 232, 76, 192, 144
    frames = []
    while len(frames) < 140:
        ok, f = cap.read()
        if not ok:
            break
        f = cv2.resize(f, (960, 540), interpolation=cv2.INTER_LINEAR)
        g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
        frames.append(g[y0:y0 + h, x0:x0 + w].copy())
    fr = np.stack(frames)
    run_detector_case("clip_192x144_n25", fr, np.ones((h, w), np.uint8), 25, 25, NORMAL, dy=True)


def case_nms():
    rng = np.random.default_rng(7)
    sets, outs, probs = [], [], []
    for k in range(60):
        N = int(rng.integers(1, 16)) if k < 45 else int(rng.integers(17, 120))
        c = rng.integers(20, 300, (N, 2))
        d = rng.integers(-40, 41, (N, 2))
        d[(d == 0).all(1)] = 3
        if k % 3 == 0:  # clustered, so that absorption actually happens
            c = c[:1] + rng.integers(-6, 7, (N, 2))
        lines = np.concatenate([c, c + d], 1).astype(np.int32)
        o, p = lineset_nms(lines)
        sets.append(lines); outs.append(np.asarray(o).reshape(-1, 4)); probs.append(np.asarray(p, np.float64).reshape(-1, 1))
    a, ao = ragged(sets, 4, np.int32)
    b, bo = ragged(outs, 4, np.int32)
    c, _ = ragged(probs, 1, np.float64)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), lines=a, lines_offs=ao, out=b, out_offs=bo,
                        prob=c[:, 0])
    print("nms: 60 sets")


def case_sliding_window():
    rng = np.random.default_rng(11)
    n, shape, T = 4, (5, 7), 11
    sw = SlidingWindow(n, shape, np.uint8, force_int=True)
    xs = rng.integers(0, 256, (T,) + shape, dtype=np.uint8)
    sw_std = SlidingWindow(n, shape, np.uint8, force_int=True, calc_std=True)  # utils.py:309-321, integer branch
    means, maxs, sums, lens, rings, stds = [], [], [], [], [], []
    for t in range(T):
        sw.update(xs[t])
        sw_std.update(xs[t])
        means.append(sw.mean.copy()); maxs.append(sw.max.copy()); sums.append(sw.sum.copy()); lens.append(sw.length)
        rings.append(sw.sliding_window.copy()); stds.append(float(sw_std.std))
    e = EMA(momentum=1 - 2 / 60, warmup_speed=25)
    vals = rng.uniform(0.2, 3, 40)
    ev = []
    for v in vals:
        e.update(v); ev.append(float(e.cur_value))
    # ROI selection on a few masks
    rois, masks = [], []
    for k in range(4):
        H, W = 120 + 17 * k, 200 + 31 * k
        m = np.ones((H, W), np.uint8)
        if k >= 1:
            m[H // 2 - 5:, :] = 0
        if k >= 2:
            m[: H // 4, : W // 2] = 0
        if k == 3:
            m[:, :] = 1; m[H // 2:H // 2 + 4, W // 2:W // 2 + 9] = 0
        s = SNR_SW(n=5, mask=m, est_snr=True, est_area=0.1 + 0.05 * k, nz_interval=2)
        rois.append(np.array(s.std_roi, np.int64))
        masks.append(np.packbits(m, axis=None))
    np.savez_compressed(os.path.join(HERE, "sliding_window.npz"), xs=xs, n=n, mean=np.stack(means),
                        max=np.stack(maxs), sum=np.stack(sums), length=np.array(lens),
                        ring=np.stack(rings), std=np.array(stds),
                        ema_in=vals, ema_out=np.array(ev), ema_momentum=1 - 2 / 60, ema_warmup=25,
                        rois=np.stack(rois), roi_mask_shapes=np.array([(120 + 17 * k, 200 + 31 * k) for k in range(4)]),
                        roi_areas=np.array([0.1 + 0.05 * k for k in range(4)]),
                        **{f"roi_mask{k}": masks[k] for k in range(4)})
    print("sliding_window: ok", rois)


def case_hough():
    """cv2.HoughLinesP itself (the un-vendored third-party call, opencv-python 4.13.0 here)."""
    rng = np.random.default_rng(5)
    masks, params, outs = [], [], []
    for k in range(40):
        H, W = int(rng.integers(40, 200)), int(rng.integers(40, 260))
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            p1 = (int(rng.integers(0, W)), int(rng.integers(0, H)))
            p2 = (int(rng.integers(0, W)), int(rng.integers(0, H)))
            cv2.line(m, p1, p2, 255, int(rng.integers(1, 4)))
        dens = [0, 0.002, 0.01, 0.05][k % 4]
        m[rng.random((H, W)) < dens] = 255
        if k % 7 == 0:
            m[rng.random((H, W)) < 0.3] = 0
        thr, ml = int(rng.integers(5, 15)), int(rng.integers(3, 20))
        gap = float(rng.uniform(0, 10)) if k % 2 else float(rng.integers(0, 10)) + 0.5
        r = cv2.HoughLinesP(m, 1, np.pi / 180 * 1, thr, minLineLength=ml, maxLineGap=gap)
        r = np.zeros((0, 4), np.int32) if r is None else r[:, 0, :]
        masks.append((np.packbits(m > 0, axis=None), H, W)); params.append((thr, ml, gap)); outs.append(r)
    o, oo = ragged(outs, 4, np.int32)
    np.savez_compressed(os.path.join(HERE, "hough.npz"), shapes=np.array([(h, w) for _, h, w in masks]),
                        params=np.array(params, np.float64), out=o, out_offs=oo,
                        **{f"mask{k}": masks[k][0] for k in range(40)})
    print("hough: 40 masks, lines per mask", [len(x) for x in outs])


def run_classic_case(name, frames, mask, fps, cfg_tuple, hough):
    """Real MetLib.Detector.ClassicDetector (Detector.py:245-299) trajectory."""
    from MetLib.Detector import ClassicDetector
    adaptive, init_value, sens, area, interval = cfg_tuple
    cfg = BinaryCfg(BinaryCoreCfg(adaptive, init_value, sens, area, interval), HoughLineCfg(*hough), DynamicCfg(False, 5))
    det = ClassicDetector(window_sec=1.0, fps=fps, mask=mask, num_cls=10, cfg=cfg, logger=BaseMetLog())
    assert det.stack_maxsize == 4
    T, H, W = frames.shape
    thr, thrf, snr, raw, cls0 = [], [], [], [], []
    masks = np.zeros((T, (H * W + 7) // 8), np.uint8)
    for t in range(T):
        det.update(frames[t])
        lines, cp = det.detect()
        thr.append(int(det.bi_threshold)); thrf.append(float(det.bi_threshold_float)); snr.append(float(det.stack.snr))
        lines = np.asarray(lines, np.int32).reshape(-1, 4)
        raw.append(lines)
        cp = np.asarray(cp, np.float64).reshape(-1, 10)
        assert len(cp) == len(lines) and (len(cp) == 0 or (np.all(cp[:, 0] == 1) and np.all(cp[:, 1:] == 0)))
        if t >= 3:
            # the reference keeps `dst` local to detect(): recompute it here from the detector's own ring
            # and threshold with the same cv2 calls (Detector.py:268-281) to record the mask as well
            sw, ci = det.stack.sliding_window, det.stack.cur_index
            d23 = cv2.threshold(cv2.absdiff(sw[ci - 1], sw[ci]), det.bi_threshold, 255, cv2.THRESH_BINARY)[1]
            d23 = 255 - cv2.dilate(d23, det.cv_op)
            dd = cv2.absdiff(cv2.bitwise_and(d23, sw[ci - 3]), cv2.bitwise_and(d23, sw[ci - 2]))
            dd = cv2.dilate(cv2.threshold(dd, det.bi_threshold, 255, cv2.THRESH_BINARY)[1], det.cv_op)
            lp = cv2.HoughLinesP(dd, rho=1, theta=np.pi / 180, threshold=hough[0], minLineLength=hough[1], maxLineGap=hough[2])
            assert np.array_equal(lines, np.zeros((0, 4), np.int32) if lp is None else lp[:, 0, :])
            masks[t] = np.packbits(dd.reshape(-1) > 0)
    r, ro = ragged(raw, 4, np.int32)
    path = os.path.join(HERE, f"classic_{name}.npz")
    np.savez_compressed(path, frames=frames, mask=mask, fps=fps, cfg=np.array([adaptive, init_value, area, interval], np.float64),
                        sens=sens, hough=np.array(hough), bi_threshold=np.array(thr), bi_threshold_float=np.array(thrf),
                        snr=np.array(snr), dst_bits=masks, raw_lines=r, raw_offs=ro)
    print("classic", name, frames.shape, "thr", sorted(set(thr)), "frames with lines", sum(len(x) > 0 for x in raw),
          "max lines", max(len(x) for x in raw))


def case_classic():
    W, H, FPS, T = 320, 240, 30, 50
    frames = synth.make_stream(T, W, H, FPS, speed_scale=3.0, thickness=2)
    run_classic_case("synth_320x240", frames, np.ones((H, W), np.uint8), FPS, (True, 7, "normal", 0.1, 2), (10, 10, 10))
    W, H, FPS, T = 203, 157, 25, 40
    frames = synth.make_stream(T, W, H, FPS, speed_scale=2.0, thickness=2, sigma=3.0)
    mask = np.ones((H, W), np.uint8); mask[:30, :50] = 0
    run_classic_case("odd_203x157_mask", frames * mask[None], mask, FPS, (True, 7, "high", 0.2, 1), (8, 8, 3))
    rng = np.random.default_rng(3)
    W, H, T = 256, 160, 12
    frames = rng.integers(0, 50, (T, H, W)).astype(np.uint8)
    run_classic_case("dense_256x160_fixed", frames, np.ones((H, W), np.uint8), 30, (False, 20, "normal", 0.1, 2), (10, 10, 10))


def case_gauss_stack():
    """FastGaussianContainer (MetLib/stacker.py:52-59) on colour frames; second case overflows uint16 sums."""
    from MetLib.stacker import FastGaussianContainer
    rng = np.random.default_rng(11)
    out = {}
    for name, T, shape, lo, hi in [("rgb40", 40, (36, 50, 3), 0, 256), ("wrap300", 300, (9, 13), 200, 256)]:
        frames = rng.integers(lo, hi, (T,) + shape, dtype=np.uint8)
        box = FastGaussianContainer()
        for f in frames:
            box.append(f)
        g = box.container
        with np.errstate(all="ignore"):
            mu, var = g.mu, g.var
        out.update({f"{name}_frames": frames, f"{name}_sum": g.sum_mu, f"{name}_sq": g.square_sum, f"{name}_n": g.n,
                    f"{name}_mu": mu, f"{name}_var": var})
        print("gauss", name, frames.shape, g.sum_mu.dtype, g.square_sum.dtype, g.n.dtype, int(g.sum_mu.max()), int(g.n.flat[0]))
    np.savez_compressed(os.path.join(HERE, "gauss_stack.npz"), names=np.array(["rgb40", "wrap300"]), **out)


def case_mfnr():
    """mfnr_mix_stacker (MetLib/stacker.py:296-403) of the live reference through a loader stub, connect_lines off,
    background algorithms "mean" and "sigma-clipping"."""
    from MetLib.stacker import mfnr_mix_stacker
    from MetLib.metstruct import ConnectParam, DenoiseOption, MFNRDenoiseParam, SimpleDenoiseParam

    class Loader:  # the protocol _batch_stacker drives (stacker.py:146-175)
        def __init__(self, frames):
            self.frames, self.i = frames, 0
            self.iterations = len(frames)

        def reset(self, start_frame=None, end_frame=None):
            self.i = 0

        def start(self):
            self.i = 0

        def pop(self):
            f = self.frames[self.i]; self.i += 1
            return f

        def stop(self):
            pass

    rng = np.random.default_rng(21)
    out = {}
    for name, T, H, W in [("clip24", 24, 72, 104), ("clip50", 50, 48, 64), ("clip13", 13, 40, 56)]:
        base = rng.integers(15, 70, (H, W, 3))
        frames = np.clip(base[None] + rng.normal(0, 4.0, (T, H, W, 3)), 0, 255).astype(np.uint8)
        for t in range(T):  # a moving streak, a saturated lamp, a hot pixel that flickers
            x = 5 + 3 * t
            if x + 8 < W:
                frames[t, H // 2 + t // 3, x:x + 8] = (230, 240, 250)
            frames[t, 6:10, 8:12] = 255
            if t % 7 == 0:
                frames[t, H - 9, W - 12] = (90, 200, 120)
        out[f"{name}_frames"] = frames
        for algo in ("mean", "sigma-clipping", "median", "med-of-med"):
            cfg = DenoiseOption(switch=True, highlight_preserve=0.9, algorithm="mfnr-mix", blur_ksize=31,
                                connect_lines=ConnectParam(switch=False, ksize_multiplier=1.5, gamma=1.0, threshold=30),
                                simple_param=SimpleDenoiseParam(10, 20, 10, 15, 6),
                                mfnr_param=MFNRDenoiseParam(bg_algorithm=algo, sigma_high=3.0, sigma_low=3.0, bg_fix_factor=1.5))
            with np.errstate(all="ignore"):
                mix = mfnr_mix_stacker(Loader(list(frames)), cfg, None, None, BaseMetLog())
            assert mix is not None and mix.dtype == np.uint8 and mix.shape == (H, W, 3)
            out[f"{name}_{algo}"] = mix
            print("mfnr", name, algo, mix.shape, int(mix.mean() * 1000) / 1000, int((mix != frames.max(0)).sum()))
    np.savez_compressed(os.path.join(HERE, "mfnr.npz"), names=np.array(["clip24", "clip50", "clip13"]), **out)


def case_preproc():
    """Loader preprocessing by the reference's own Transform (MetLib/imgproc.py:70-139) and
    MergeFunction.max (MetLib/utils.py:203-204): resize -> BGR2GRAY -> mask, exp_frame merge."""
    from MetLib.imgproc import Transform
    from MetLib.utils import MergeFunction
    rng = np.random.default_rng(77)
    out = {}
    cases = [("down4_rgb", 8, (192, 108, 3), (48, 27), True, 2), ("ratio_rgb", 5, (150, 90, 3), (64, 37), True, 1),
             ("up_gray", 4, (40, 30, 1), (96, 54), False, 3), ("same_rgb", 3, (64, 32, 3), (64, 32), True, 1)]
    for name, T, (W0, H0, C), (W, H), gray, exp in cases:
        base = rng.integers(0, 256, (H0, W0, C), dtype=np.uint8)
        frames = np.stack([np.clip(base.astype(int) + rng.integers(-40, 40, base.shape), 0, 255).astype(np.uint8)
                           for _ in range(T)])
        if C == 1:
            frames = frames[..., 0]
        mask = (rng.random((H, W)) > 0.2).astype(np.uint8)
        tr = Transform()
        if (W0, H0) != (W, H):
            tr.opencv_resize([W, H])
        if gray:
            tr.opencv_BGR2GRAY()
        tr.mask_with(mask)
        res = []
        for s in range(0, T, exp):
            group = [tr.exec_transform(f) for f in frames[s:s + exp]]
            res.append(group[0] if len(group) == 1 else MergeFunction.max(group))
        out[f"{name}_frames"] = frames
        out[f"{name}_mask"] = mask
        out[f"{name}_out"] = np.stack(res)
        out[f"{name}_cfg"] = np.array([W, H, int(gray), exp])
        print("preproc", name, frames.shape, "->", out[f"{name}_out"].shape)
    np.savez_compressed(os.path.join(HERE, "preproc.npz"), names=np.array([c[0] for c in cases]), **out)


def _clip_config1_frames(W, H):
    """The bundled clip exactly as MetDetPy.detect_video feeds it to the detector with config/m3det_normal.json
    (BASELINE config 1): cv2/FFmpeg decode -> resize (W, H) INTER_LINEAR -> BGR2GRAY -> x mask (test/mask-east.jpg
    through the reference's own fileio.load_mask) -> MergeFunction.max over exp_frame = 4 decoded frames
    (videoloader.py:300-308, :388; utils.py:203-204).  eq_fps = 25 / 4 = 6.25, window n = int(1 * 6.25) = 6."""
    from MetLib.fileio import load_mask
    from MetLib.utils import MergeFunction
    mask = load_mask("/root/reference/test/mask-east.jpg", [W, H], grayscale=True)
    assert mask.shape == (H, W) and set(np.unique(mask)) <= {0, 1}
    cap = cv2.VideoCapture("/root/reference/test/20220413Red.mp4")
    merged, group = [], []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        f = cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR)
        g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) * mask
        group.append(g)
        if len(group) == 4:
            merged.append(MergeFunction.max(group)); group = []
    if group:
        merged.append(MergeFunction.max(group))
    return np.stack(merged).astype(np.uint8), mask


def case_clip_config1():
    """BASELINE config 1 through the real M3Detector: the whole clip at 480x270 (80 detector iterations) and the
    stretch around the annotated meteor (test/20220413_annotation.json: 2.4 s .. 4.4 s) at the runtime size 960x540."""
    fr, mask = _clip_config1_frames(480, 270)
    run_detector_case("clip_cfg1_480x270_n6", fr, mask, 6, 6.25, NORMAL, dy=True)
    fr, mask = _clip_config1_frames(960, 540)
    run_detector_case("clip_cfg1_960x540_n6_range", fr[8:30], mask, 6, 6.25, NORMAL, dy=True)


def case_masks():
    """test/mask-east.jpg through fileio.load_mask (fileio.py:250-292) at the sizes the benchmark and the tests use,
    bit-packed."""
    from MetLib.fileio import load_mask
    for W, H in ((3840, 2160), (960, 540), (480, 270)):
        m = load_mask("/root/reference/test/mask-east.jpg", [W, H], grayscale=True)
        assert m.shape == (H, W) and m.dtype == np.uint8 and set(np.unique(m)) <= {0, 1}
        path = os.path.join(HERE, f"mask_east_{W}x{H}.npz")
        np.savez_compressed(path, bits=np.packbits(m, axis=None), shape=np.array([H, W]))
        print(f"mask_east {W}x{H}: open share {m.mean():.4f} -> {os.path.getsize(path) / 1e3:.1f} KB")


CASES = dict(mfnr=case_mfnr, clip_cfg1=case_clip_config1, masks=case_masks, preproc=case_preproc, gauss=case_gauss_stack, classic=case_classic, synth_small=case_synth_small, synth_dy_mask=case_synth_dy_mask, odd=case_odd_size,
             dense=case_fixed_thr_dense, low=case_low_sens, clip=case_real_clip, nms=case_nms,
             sw=case_sliding_window, hough=case_hough)

if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    print("cv2", cv2.__version__, "numpy", np.__version__)
    for c in which:
        CASES[c]()
