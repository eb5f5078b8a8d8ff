"""CPU: pin the oracle (oracle/m3_oracle.py + oracle/ppht.c) against the golden vectors produced
by the live reference (tests/golden/make_golden.py).  Both backends of the oracle are checked:
"cv2" (reference call sites) and "numpy" (pure restatements of the five cv2 functions)."""
import os

import numpy as np
import pytest

from conftest import (DET_CASES, GOLDEN, assert_nms_equivalent, has_len2_ties, load_det_case,
                      ragged_get)
from oracle import m3_oracle as O


@pytest.mark.parametrize("backend", ["cv2", "numpy"])
@pytest.mark.parametrize("name", DET_CASES)
def test_detector_trajectory(name, backend):
    if backend == "cv2" and O.cv2 is None:
        pytest.skip("cv2 not importable")
    g = load_det_case(name)
    det = O.M3DetectorOracle(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, backend=backend,
                             **g["cfg"])
    assert det.stack_maxsize == g["n"]
    assert tuple(det.stack.std_roi) == tuple(g["std_roi"])
    assert int(det.mask_area) == int(g["mask_area"])
    T = len(g["frames"])
    for t in range(T):
        det.update(g["frames"][t])
        lines, cls = det.detect()
        assert det.bi_threshold == g["bi_threshold"][t], t
        assert det.bi_threshold_float == pytest.approx(g["bi_threshold_float"][t], rel=1e-12), t
        assert float(det.stack.snr) == pytest.approx(g["snr"][t], rel=1e-12, abs=0), t
        assert np.array_equal(det.dst, g["dst"][t]), f"dst differs at frame {t}"
        assert det.dst_sum == g["dst_sum"][t]
        raw = ragged_get(g["raw_lines"], g["raw_offs"], t)
        assert det.lines_num == g["lines_num"][t], t
        assert np.array_equal(np.asarray(det.linesp_ext).reshape(-1, 4), raw), t
        ref = ragged_get(g["nms_lines"], g["nms_offs"], t)
        refc = ragged_get(g["cls_pred"], g["nms_offs"], t)
        assert_nms_equivalent(lines, np.asarray(cls).reshape(-1, 10)[:, -1], ref, refc[:, -1], raw, t)


def test_cv_restatements_match_cv2():
    if O.cv2 is None:
        pytest.skip("cv2 not importable")
    cv2 = O.cv2
    rng = np.random.default_rng(0)
    k = np.ones((3, 3), np.uint8)
    for (H, W) in [(1, 1), (1, 7), (6, 1), (2, 2), (3, 5), (37, 53), (64, 64)]:
        a = rng.integers(0, 256, (H, W), dtype=np.uint8)
        assert np.array_equal(O.median3(a), cv2.medianBlur(a, 3))
        for thr in (0, 3, 100, 254, 255):
            assert np.array_equal(O.threshold_binary(a, thr),
                                  cv2.threshold(a, thr, 255, cv2.THRESH_BINARY)[1])
        b = np.where(rng.random((H, W)) < 0.3, 255, 0).astype(np.uint8)
        assert np.array_equal(O.close3(b), cv2.morphologyEx(b, cv2.MORPH_CLOSE, k))
        assert np.array_equal(O.erode3(b), cv2.erode(b, k))
        assert np.array_equal(O.dilate3(b), cv2.dilate(b, k))
        m = (b > 0).astype(np.uint8)
        assert np.array_equal(O.erode3(m), cv2.erode(m, k))


def test_ppht_golden_fixtures():
    g = np.load(os.path.join(GOLDEN, "hough.npz"))
    for k in range(len(g["shapes"])):
        H, W = (int(v) for v in g["shapes"][k])
        m = np.unpackbits(g[f"mask{k}"])[:H * W].reshape(H, W) * np.uint8(255)
        thr, ml, gap = g["params"][k]
        out = O.hough_lines_p(m, int(thr), ml, gap)
        ref = ragged_get(g["out"], g["out_offs"], k)
        assert np.array_equal(out, ref), k


def test_ppht_matches_live_cv2():
    if O.cv2 is None:
        pytest.skip("cv2 not importable")
    cv2 = O.cv2
    rng = np.random.default_rng(42)
    for k in range(60):
        H, W = int(rng.integers(30, 300)), int(rng.integers(30, 400))
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            cv2.line(m, (int(rng.integers(0, W)), int(rng.integers(0, H))),
                     (int(rng.integers(0, W)), int(rng.integers(0, H))), 255, int(rng.integers(1, 4)))
        m[rng.random((H, W)) < [0, 0.003, 0.02, 0.08][k % 4]] = 255
        thr, ml, gap = int(rng.integers(4, 14)), int(rng.integers(2, 20)), float(rng.uniform(0, 10))
        r = cv2.HoughLinesP(m, 1, np.pi / 180, thr, minLineLength=ml, maxLineGap=gap)
        r = np.zeros((0, 4), np.int32) if r is None else r[:, 0, :]
        assert np.array_equal(O.hough_lines_p(m, thr, ml, gap), r), k
    # edge cases: empty image, single pixel, full image
    for m in (np.zeros((20, 30), np.uint8), np.full((9, 9), 255, np.uint8)):
        r = cv2.HoughLinesP(m, 1, np.pi / 180, 5, minLineLength=3, maxLineGap=1)
        r = np.zeros((0, 4), np.int32) if r is None else r[:, 0, :]
        assert np.array_equal(O.hough_lines_p(m, 5, 3, 1), r)


def test_nms_golden():
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    for k in range(len(g["lines_offs"]) - 1):
        lines = ragged_get(g["lines"], g["lines_offs"], k)
        ref = ragged_get(g["out"], g["out_offs"], k)
        refp = ragged_get(g["prob"], g["out_offs"], k)
        out, p = O.lineset_nms(lines)
        assert_nms_equivalent(out, p, ref, refp, lines, k)


def test_sliding_window_ema_roi_golden():
    g = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    sw = O.SlidingWindow(int(g["n"]), g["xs"].shape[1:])
    for t, x in enumerate(g["xs"]):
        sw.update(x)
        assert sw.length == g["length"][t]
        assert np.array_equal(sw.mean, g["mean"][t]) and sw.mean.dtype == np.uint8
        assert np.array_equal(sw.max, g["max"][t])
        assert np.array_equal(sw.sum, g["sum"][t])
        assert np.array_equal(sw.sliding_window, g["ring"][t])
        assert sw.std == g["std"][t]
    e = O.EMA(float(g["ema_momentum"]), float(g["ema_warmup"]))
    for v, r in zip(g["ema_in"], g["ema_out"]):
        e.update(v)
        assert e.cur_value == r
    for k in range(4):
        H, W = (int(v) for v in g["roi_mask_shapes"][k])
        m = np.unpackbits(g[f"roi_mask{k}"])[:H * W].reshape(H, W)
        assert O.select_subarea(m, float(g["roi_areas"][k])) == tuple(int(v) for v in g["rois"][k])


def test_stack_helpers():
    rng = np.random.default_rng(1)
    fr = rng.integers(0, 256, (7, 5, 6, 3), dtype=np.uint8)
    assert np.array_equal(O.max_stack(list(fr)), fr.max(0))
    assert O.max_stack([]) is None
    assert np.array_equal(O.merge_max(list(fr[:4])), fr[:4].max(0))
