"""GPU parity at BASELINE.json's full sizes (configs 2-5), through the C ABI.

(1) against the CPU oracle (cv2 backend = the reference's own numpy+cv2 calls) on a bounded number of
    frames of the SURVEY App. E synthetic stream -- masks, thresholds and segments must be identical;
(2) size-independent properties on longer runs: the streaming kernels and the generic per-frame
    kernel (two independent implementations) agree bit for bit; splitting the stream into different
    batch sizes changes nothing; `on-pixel count == popcount(dst)`; dst is {0,255}; zero-copy device
    input == host input."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_nms_equivalent

pytestmark = pytest.mark.gpu


def _cfg(dy=True):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    return BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(dy, 5))


def _real_mask(H, W):
    """test/mask-east.jpg of the reference through fileio.load_mask (MetLib/fileio.py:250-292) at this size, committed
    bit-packed by tests/golden/make_golden.py (the reference tree does not exist on the GPU box)."""
    z = np.load(os.path.join(GOLDEN, f"mask_east_{W}x{H}.npz"))
    return np.unpackbits(z["bits"])[:H * W].reshape(H, W).astype(np.uint8)


@pytest.mark.parametrize("name,W,H,fps,n,dy,masked,T", [
    ("config2_1080p_n5", 1920, 1080, 30, 5, False, False, 40),
    ("config3_4k_n30", 3840, 2160, 30, 30, True, False, 48),
    ("config4_4k60_n60_mask", 3840, 2160, 60, 60, True, True, 76),
    ("config5_8k_n30", 7680, 4320, 30, 30, True, False, 40),
])
def test_baseline_config_against_oracle(name, W, H, fps, n, dy, masked, T):
    from metdetpy_b200 import synth
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    frames = synth.make_stream(T, W, H, fps)
    mask = _real_mask(H, W) if masked else np.ones((H, W), np.uint8)
    ref = O.M3DetectorOracle(n / fps + 1e-9, fps, mask, 10, adaptive=True, init_value=7, sensitivity="normal",
                             area=0.1, interval=2, hough=(10, 10, 10), dy_mask=dy,
                             backend="cv2" if O.cv2 is not None else "numpy")
    # the loader's mask_with (imgproc.py:96-101) happens on the device (apply_mask) for the masked config
    det = M3Detector(n / fps + 1e-9, fps, mask, 10, _cfg(dy), None, max_batch=16, apply_mask=masked)
    assert tuple(det.stack.std_roi) == tuple(ref.stack.std_roi)
    nz_frames = 0
    for s in range(0, T, 16):
        res, dst = det.detect_many(frames[s:s + 16], return_dst=True)
        for i, (lines, cls) in enumerate(res):
            t = s + i
            f = frames[t] * mask if masked else frames[t]
            ref.update(f)
            rl, rc = ref.detect()
            info = det.last_infos[i]
            assert info["bi_threshold"] == ref.bi_threshold, (t, info["bi_threshold"], ref.bi_threshold)
            assert info["snr"] == pytest.approx(float(ref.stack.snr), rel=1e-12, abs=0), t
            assert np.array_equal(dst[i], ref.dst), (t, int(np.count_nonzero(dst[i] != ref.dst)))
            assert info["dst_sum"] == ref.dst_sum and info["gap"] == ref.gap, t
            assert info["lines_num"] == ref.lines_num, t
            raw = np.asarray(ref.linesp_ext).reshape(-1, 4)
            assert np.array_equal(det.last_raw[i].reshape(-1, 4), raw), t
            assert_nms_equivalent(lines, np.asarray(cls).reshape(-1, 10)[:, -1], rl,
                                  np.asarray(rc).reshape(-1, 10)[:, -1], raw, t)
            nz_frames += int(info["n_on"] > 0)
    assert nz_frames > 0, "the stream never lit a pixel: generator or thresholds changed"
    det.close()


def test_4k_streaming_vs_generic_kernel_and_batch_split_invariance():
    import torch
    from metdetpy_b200 import synth
    from metdetpy_b200.detector import M3Detector
    W, H, fps, n, T = 3840, 2160, 30, 30, 96
    dev = torch.device("cuda", 0)
    xd = synth.make_stream_device(T, W, H, fps, dev)
    torch.cuda.synchronize()
    frames = xd.cpu().numpy()
    mask = np.ones((H, W), np.uint8)

    def run(batches, stream_kernel, on_device):
        det = M3Detector(n / fps + 1e-9, fps, mask, 10, _cfg(True), None, max_batch=max(batches))
        det._eng.set_option("stream_kernel", stream_kernel)
        out, s = [], 0
        for b in batches:
            if on_device:
                res, dst = det.detect_many((xd[s:s + b].data_ptr(), b), on_device=True, return_dst=True)
            else:
                res, dst = det.detect_many(frames[s:s + b], return_dst=True)
            for i in range(b):
                info = det.last_infos[i]
                assert set(np.unique(dst[i])) <= {0, 255}
                assert int(np.count_nonzero(dst[i])) == info["n_on"]
                out.append((dst[i].copy(), info["bi_threshold"], info["lines_num"], det.last_raw[i].copy(),
                            np.asarray(res[i][0]).reshape(-1, 4).copy()))
            s += b
        det.close()
        return out

    a = run([96], 1, True)                 # one batch, streaming kernels, zero-copy device input
    b = run([7, 25, 64], 1, False)         # ragged batches, host input
    c = run([48, 48], 0, False)            # generic per-frame kernel
    lit = 0
    for t in range(T):
        for other in (b, c):
            assert np.array_equal(a[t][0], other[t][0]), t
            assert a[t][1] == other[t][1] and a[t][2] == other[t][2], t
            assert np.array_equal(a[t][3], other[t][3]) and np.array_equal(a[t][4], other[t][4]), t
        lit += a[t][2] > 0
    assert lit > 10


def test_4k_two_far_apart_objects_stay_on_chip_and_equal_cv2():
    """Two streaks far apart in one window (plus a dense-noise frame): per angle their points project onto two rho
    clusters, so the shared-memory PPHT tier keeps TWO intervals per row instead of handing the frame to the
    global-memory tier.  Raw segments must equal cv2.HoughLinesP on the device's own masks, and the tier statistics
    must show the frames on chip."""
    import cv2
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    W, H, n, T = 3840, 2160, 5, 24
    rng = np.random.default_rng(21)
    frames = np.clip(48 + rng.normal(0, 2.0, (T, H, W)), 0, 255).astype(np.uint8)
    for t in range(4, T):  # two objects, ~3000 px apart, moving in different directions
        for (x0, y0, dx, dy) in ((300, 250, 14, 5), (3400, 1900, -9, -12)):
            m = np.zeros((H, W), np.uint8)
            p1 = (int(x0 + dx * (t - 4)), int(y0 + dy * (t - 4)))
            p2 = (int(x0 + dx * (t - 3)), int(y0 + dy * (t - 3)))
            cv2.line(m, p1, p2, 70, 3, cv2.LINE_AA)
            frames[t] = np.clip(frames[t].astype(np.int32) + m, 0, 255).astype(np.uint8)
    mask = np.ones((H, W), np.uint8)
    cfg = BinaryCfg(BinaryCoreCfg(False, 12, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(True, 5))
    det = M3Detector(n / 30 + 1e-9, 30, mask, 10, cfg, None, max_batch=T)
    res, dst = det.detect_many(frames, return_dst=True)
    both = 0
    for t in range(T):
        info = det.last_infos[t]
        ref = cv2.HoughLinesP(dst[t], 1, np.pi / 180, 10, minLineLength=10, maxLineGap=float(info["gap"]))
        ref = np.zeros((0, 4), np.int32) if ref is None else ref.reshape(-1, 4)
        assert info["lines_num"] == len(ref), (t, info["lines_num"], len(ref))
        assert np.array_equal(det.last_raw[t].reshape(-1, 4), ref), t
        ys, xs = np.nonzero(dst[t])
        both += int(len(xs) > 0 and xs.min() < 1500 and xs.max() > 2500)
    assert both >= 10, "the two objects were not in the masks together"
    on_chip = det._eng.info("hough_tier1a") + det._eng.info("hough_tier1b")
    assert det._eng.info("hough_tier2") == 0 and det._eng.info("hough_tier3") == 0, "a frame left the shared-memory tiers"
    assert on_chip >= both
    det.close()
