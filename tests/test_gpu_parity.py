"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
(1) the golden vectors of the live reference and (2) the CPU oracle on seeded inputs.
Bars: bi_threshold, dst mask, on-pixel count, dst_sum, gap, raw Hough segments: bit-exact.
snr / bi_threshold_float: relative 1e-12 (float64 EMA of a std computed from exact integer sums).
NMS lines: exact; when two raw segments tie in length and there are more than 16 of them, numpy's
unstable argsort makes the reference itself implementation-defined, so sets are compared."""
import os

import numpy as np
import pytest

from conftest import (DET_CASES, GOLDEN, assert_nms_equivalent, has_len2_ties, load_det_case,
                      ragged_get)

pytestmark = pytest.mark.gpu


def _cfg(c):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    return BinaryCfg(BinaryCoreCfg(c["adaptive"], c["init_value"], c["sensitivity"], c["area"], c["interval"]),
                     HoughLineCfg(*c["hough"]), DynamicCfg(c["dy_mask"], 5))


def _check_frame(det, g, t, lines, cls, dst, info=None):
    thr = det.bi_threshold if info is None else info["bi_threshold"]
    thrf = det.bi_threshold_float if info is None else info["bi_threshold_float"]
    snr = det.stack.snr if info is None else info["snr"]
    dsum = det.dst_sum if info is None else info["dst_sum"]
    nraw = det.lines_num if info is None else info["lines_num"]
    assert thr == g["bi_threshold"][t], (t, thr, g["bi_threshold"][t])
    assert thrf == pytest.approx(g["bi_threshold_float"][t], rel=1e-12), t
    assert snr == pytest.approx(g["snr"][t], rel=1e-12, abs=0), t
    assert np.array_equal(dst, g["dst"][t]), f"dst differs at frame {t}: {np.count_nonzero(dst != g['dst'][t])} px"
    assert dsum == g["dst_sum"][t], t
    assert nraw == g["lines_num"][t], (t, nraw, g["lines_num"][t])
    ref = ragged_get(g["nms_lines"], g["nms_offs"], t)
    raw = ragged_get(g["raw_lines"], g["raw_offs"], t)
    refc = ragged_get(g["cls_pred"], g["nms_offs"], t)
    assert_nms_equivalent(lines, np.asarray(cls).reshape(-1, 10)[:, -1], ref, refc[:, -1], raw, t)
    assert len(lines) == 0 or (lines.dtype == np.int32 and cls.dtype == np.float64)


@pytest.mark.parametrize("name", DET_CASES)
def test_per_frame_api_matches_reference_golden(name):
    from metdetpy_b200.detector import M3Detector
    g = load_det_case(name)
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None)
    assert det.stack_maxsize == g["n"]
    assert tuple(det.stack.std_roi) == tuple(g["std_roi"])
    assert int(det.mask_area) == int(g["mask_area"])
    for t in range(len(g["frames"])):
        det.update(g["frames"][t])
        lines, cls = det.detect()
        _check_frame(det, g, t, lines, cls, det.dst)
        raw = ragged_get(g["raw_lines"], g["raw_offs"], t)
        assert np.array_equal(np.asarray(det.linesp_ext).reshape(-1, 4), raw), t
    det.close()


@pytest.mark.parametrize("mode", ["generic", "stream", "stream_dense_dst", "stream_strip_act", "stream_temporal_v2",
                                  "stream_subblocks"])
@pytest.mark.parametrize("batch", [1, 7, 32])
@pytest.mark.parametrize("name", DET_CASES)
def test_batched_api_matches_reference_golden(name, batch, mode):
    """generic = one fused launch per frame; stream = temporal3 (register ring; temporal2 for single-frame
    batches) + act4/act + sparse dst; the extra modes force the full-scan dst kernel (list-overflow path), the
    warp-strip act kernel and the second-generation temporal kernel."""
    from metdetpy_b200.detector import M3Detector
    g = load_det_case(name)
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None,
                     max_batch=batch)
    stream_kernel = int(mode != "generic")
    W = g["frames"].shape[2]
    det._eng.set_option("stream_kernel", stream_kernel)
    det._eng.set_option("force_dense", int(mode == "stream_dense_dst"))
    det._eng.set_option("force_strip", int(mode == "stream_strip_act"))
    det._eng.set_option("temporal_version", {"stream_temporal_v2": 2, "stream_subblocks": 2}.get(mode, 3))
    if mode == "stream_subblocks":  # sub-blocked van Herk (long windows use it by default): force a split of n
        k = next((k for k in (5, 4, 3, 2) if g["n"] % k == 0 and g["n"] // k >= 2), 0)
        if not k or W % 32 or not 2 <= g["n"] <= 128:
            pytest.skip("window not divisible / streaming path not used")
        det._eng.set_option("temporal_kdiv", k)
    T = len(g["frames"])
    W = g["frames"].shape[2]
    for s in range(0, T, batch):
        res, dst = det.detect_many(g["frames"][s:s + batch], return_dst=True)
        nb = len(res)
        streamed = stream_kernel and W % 32 == 0 and 2 <= g["n"] <= 128
        assert det._eng.fused_time()[1] == (4 if streamed else nb)  # temporal+act+dst(sparse,dense) vs one launch per frame
        for i, (lines, cls) in enumerate(res):
            _check_frame(det, g, s + i, lines, cls, dst[i], det.last_infos[i])
            raw = ragged_get(g["raw_lines"], g["raw_offs"], s + i)
            assert np.array_equal(det.last_raw[i].reshape(-1, 4), raw), s + i
    det.close()


@pytest.mark.parametrize("name", ["synth_384x216_n12_dyon_mask", "synth_256x160_n6_fixed3_dense"])
def test_sparse_and_dense_dst_kernels_share_the_mask_buffer(name):
    """The persistent u8 mask buffer, its 1-bit shadow and the non-zero word list must stay consistent
    when batches alternate between the list-driven dst kernel, the full-scan one and the generic
    per-frame kernel, with batch sizes that change (slots reused after being skipped)."""
    from metdetpy_b200.detector import M3Detector
    g = load_det_case(name)
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None, max_batch=9)
    T = len(g["frames"])
    s, k = 0, 0
    sizes = [9, 4, 9, 1, 7, 9, 2]
    while s < T:
        b = min(sizes[k % len(sizes)], T - s)
        det._eng.set_option("force_dense", int(k % 3 == 1))
        det._eng.set_option("stream_kernel", int(k % 5 != 3))
        res, dst = det.detect_many(g["frames"][s:s + b], return_dst=True)
        for i, (lines, cls) in enumerate(res):
            _check_frame(det, g, s + i, lines, cls, dst[i], det.last_infos[i])
        s += b
        k += 1
    det.close()


def test_apply_mask_on_device_equals_host_mask_with():
    """Transform.mask_with (imgproc.py:96-101) fused into the device loads."""
    from metdetpy_b200.detector import M3Detector
    g = load_det_case("synth_384x216_n12_dyon_mask")
    rng = np.random.default_rng(3)
    raw = g["frames"].copy()
    hole = g["mask"] == 0
    raw[:, hole] = rng.integers(0, 256, (len(raw), int(hole.sum())), dtype=np.uint8)  # junk under the mask
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None,
                     max_batch=16, apply_mask=True)
    for s in range(0, len(raw), 16):
        res, dst = det.detect_many(raw[s:s + 16], return_dst=True)
        for i, (lines, cls) in enumerate(res):
            _check_frame(det, g, s + i, lines, cls, dst[i], det.last_infos[i])


@pytest.mark.parametrize("seed,W,H,n,dy,sens", [(0, 97, 61, 4, True, "normal"), (1, 128, 64, 9, False, "high"),
                                                (2, 517, 131, 16, True, "low"), (3, 64, 300, 2, True, "normal"),
                                                (4, 33, 17, 1, True, "normal"), (5, 260, 200, 31, True, "normal")])
def test_random_streams_against_oracle(seed, W, H, n, dy, sens):
    """Seeded noise + moving bright bars, odd sizes, windows 1..31: CUDA vs CPU oracle (numpy backend)."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    rng = np.random.default_rng(seed)
    T = 3 * n + 7
    base = rng.integers(20, 60, (H, W))
    frames = np.clip(base[None] + rng.normal(0, 2.5, (T, H, W)), 0, 255).astype(np.uint8)
    for t in range(T):  # a bar sweeping through, plus a static hot blob (dy-mask food)
        x = (5 * t) % max(W - 12, 1)
        y = (3 * t) % max(H - 3, 1)
        frames[t, y:y + 2, x:x + 12] = 200
        frames[t, H // 3:H // 3 + 3, W // 4:W // 4 + 3] = 255 if (t // 2) % 2 == 0 or t > T // 2 else 30
    mask = np.ones((H, W), np.uint8)
    mask[: H // 8, : W // 5] = 0
    frames *= mask[None]
    kw = dict(adaptive=True, init_value=7, sensitivity=sens, area=0.2, interval=1, hough=(6, 6, 4), dy_mask=dy)
    ref = O.M3DetectorOracle(n / 10 + 1e-9, 10, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, sens, 0.2, 1), HoughLineCfg(6, 6, 4), DynamicCfg(dy, 5))
    det = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=11)
    det1 = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None)
    got, dsts, infos = [], [], []
    for s in range(0, T, 11):
        r, d = det.detect_many(frames[s:s + 11], return_dst=True)
        got += r; dsts += list(d); infos += list(det.last_infos)
    for t in range(T):
        ref.update(frames[t]); rl, rc = ref.detect()
        det1.update(frames[t]); l1, c1 = det1.detect()
        assert infos[t]["bi_threshold"] == ref.bi_threshold == det1.bi_threshold, t
        assert infos[t]["snr"] == pytest.approx(float(ref.stack.snr), rel=1e-12, abs=0), t
        assert np.array_equal(dsts[t], ref.dst), (t, int(np.count_nonzero(dsts[t] != ref.dst)))
        assert np.array_equal(det1.dst, ref.dst), t
        assert infos[t]["dst_sum"] == ref.dst_sum and infos[t]["gap"] == ref.gap, t
        assert infos[t]["lines_num"] == ref.lines_num, t
        raw = np.asarray(ref.linesp_ext).reshape(-1, 4)
        assert np.array_equal(np.asarray(det1.linesp_ext).reshape(-1, 4), raw), t
        assert_nms_equivalent(got[t][0], got[t][1][:, -1], rl, rc[:, -1], raw, t)
        assert_nms_equivalent(l1, c1[:, -1], rl, rc[:, -1], raw, t)
    # stack.max / stack.mean read-back (SlidingWindow.max/.mean, utils.py:288-300)
    assert np.array_equal(det1.stack.max, ref.stack.max)
    assert np.array_equal(det1.stack.mean, ref.stack.mean)
    assert np.array_equal(det1.stack.sum, ref.stack.sum)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20, 21, 24, 25, 28, 30, 32, 36, 40, 48, 50,
                               60, 64, 11, 33])
@pytest.mark.parametrize("apply_mask", [False, True])
def test_temporal3_every_window_shape_against_oracle(n, apply_mask):
    _temporal3_case(n, apply_mask, 0)


@pytest.mark.parametrize("variant", [1, 2, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 18, 19, 20, 21, 22, 23])
def test_temporal3_tuning_variants_against_oracle(variant):
    """The alternative shapes of the n = 30 window (temporal3_dispatch.cuh): register-fed (1-6) and bulk-copy-fed
    (cp.async.bulk + mbarrier stage ring, 7-12)."""
    _temporal3_case(30, variant % 2 == 0, variant)


def _temporal3_case(n, apply_mask, variant):
    """Every (U, BL, P, K) shape of temporal3_kernel (temporal3_dispatch.cuh; n = 11 and 33 have none and take
    temporal2): batches that end inside a van Herk block, start in the warm-up, and hand history over through the
    ring; device-side masking.  Mask, threshold and on-pixel count must equal the CPU oracle bit for bit."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    rng = np.random.default_rng(100 + n)
    W, H = 96, 40
    sizes = [n + 3, 2, 2 * n + 1, 7, 3 * n - 1, n, 1, n + 1]
    T = sum(sizes)
    base = rng.integers(20, 60, (H, W))
    frames = np.clip(base[None] + rng.normal(0, 2.5, (T, H, W)), 0, 255).astype(np.uint8)
    for t in range(T):
        x, y = (7 * t) % (W - 14), (3 * t) % (H - 3)
        frames[t, y:y + 2, x:x + 14] = 180 + (t % 70)
        if t % 9 == 0:
            frames[t, H // 2:H // 2 + 4, W // 2:W // 2 + 4] = 255
    mask = np.ones((H, W), np.uint8)
    mask[: H // 6, : W // 3] = 0
    masked = frames * mask[None]
    kw = dict(adaptive=True, init_value=7, sensitivity="normal", area=0.2, interval=1, hough=(6, 6, 4), dy_mask=True)
    ref = O.M3DetectorOracle(n / 10 + 1e-9, 10, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 1), HoughLineCfg(6, 6, 4), DynamicCfg(True, 5))
    det = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=max(sizes), apply_mask=apply_mask)
    det._eng.set_option("t3_variant", variant)
    s = 0
    for b in sizes:
        res, dst = det.detect_many((frames if apply_mask else masked)[s:s + b], return_dst=True)
        for i in range(b):
            ref.update(masked[s + i]); ref.detect()
            info = det.last_infos[i]
            assert info["bi_threshold"] == ref.bi_threshold, (s + i)
            assert np.array_equal(dst[i], ref.dst), (n, s + i, int(np.count_nonzero(dst[i] != ref.dst)))
            assert info["n_on"] == int(np.count_nonzero(ref.dst)), s + i
        if b > 1:  # single host frames go through the ring and the second-generation kernel
            assert det._eng.info("temporal_generation") == (2 if n in (11, 33) else 3), (n, b)
        s += b
    det.close()


def test_stack_readback_after_device_batches():
    """SNR_SW.max / .mean / .sum (utils.py:288-300) after batched calls: the ring must hold the last n frames of a
    batch, not n-1 (the window of the newest frame)."""
    from metdetpy_b200 import BinaryCfg
    from metdetpy_b200.detector import M3Detector
    rng = np.random.default_rng(5)
    n, H, W = 6, 32, 64
    frames = rng.integers(0, 256, (40, H, W), dtype=np.uint8)
    det = M3Detector(n / 10 + 1e-9, 10, np.ones((H, W), np.uint8), 10, BinaryCfg(), None, max_batch=16)
    s = 0
    for b in [16, 3, 16, 5]:
        det.detect_many(frames[s:s + b])
        s += b
        win = frames[max(0, s - n):s].astype(np.uint32)
        assert np.array_equal(det.stack.max, win.max(0).astype(np.uint8)), s
        assert np.array_equal(det.stack.sum, win.sum(0)), s
        assert np.array_equal(det.stack.mean, (win.sum(0) // min(n, s)).astype(np.uint8)), s
    det.close()


def test_three_batches_in_flight_generic_kernel():
    """submit / collect with three batches in flight on the generic per-frame kernel (width not a multiple of 32):
    ring slots and staging buffers must not be overwritten while an older batch still reads them."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    rng = np.random.default_rng(8)
    n, H, W, B = 5, 120, 203, 6
    T = 14 * B
    frames = np.clip(40 + rng.normal(0, 3, (T, H, W)), 0, 255).astype(np.uint8)
    for t in range(T):
        frames[t, (3 * t) % (H - 2):(3 * t) % (H - 2) + 2, (5 * t) % (W - 20):(5 * t) % (W - 20) + 20] = 220
    mask = np.ones((H, W), np.uint8)
    kw = dict(adaptive=True, init_value=7, sensitivity="normal", area=0.2, interval=1, hough=(6, 6, 4), dy_mask=True)
    ref = O.M3DetectorOracle(n / 10 + 1e-9, 10, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 1), HoughLineCfg(6, 6, 4), DynamicCfg(True, 5))
    det = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=B)
    want = []
    for t in range(T):
        ref.update(frames[t]); ref.detect()
        want.append((ref.bi_threshold, int(np.count_nonzero(ref.dst)), ref.lines_num))
    got = []
    nb = T // B
    for k in range(min(3, nb)):
        det.submit(frames[k * B:(k + 1) * B].ctypes.data, B, False)
    for k in range(nb):
        det.collect()
        got += [(int(i["bi_threshold"]), int(i["n_on"]), int(i["lines_num"])) for i in det.last_infos[:B]]
        if k + 3 < nb:
            det.submit(frames[(k + 3) * B:(k + 4) * B].ctypes.data, B, False)
    assert got == want
    det.close()


@pytest.mark.parametrize("thr,lo,hi", [(12, 4096, 16384), (10, 16384, 1 << 30)])
def test_dense_mask_overflow_path_and_too_many_lines(thr, lo, hi):
    """Masks beyond the shared-memory PPHT tiers: 4096 < on-pixels <= 16384 (tier 2: global-memory accumulator,
    one CTA per frame) and > 16384 (tier 3: point list in global memory); > 500 raw lines (Detector.py:358-360)."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    rng = np.random.default_rng(9)
    H, W, n, T = 240, 320, 3, 8
    frames = rng.integers(0, 40, (T, H, W)).astype(np.uint8)
    mask = np.ones((H, W), np.uint8)
    kw = dict(adaptive=False, init_value=thr, sensitivity="normal", area=0.1, interval=2, hough=(10, 10, 10), dy_mask=False)
    ref = O.M3DetectorOracle(n / 10 + 1e-9, 10, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(False, thr, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(False, 5))
    det = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=T)
    res, dst = det.detect_many(frames, return_dst=True)
    seen_overflow = seen_toomuch = False
    for t in range(T):
        ref.update(frames[t]); rl, rc = ref.detect()
        assert np.array_equal(dst[t], ref.dst), t
        info = det.last_infos[t]
        assert info["lines_num"] == ref.lines_num, (t, info["lines_num"], ref.lines_num)
        seen_overflow |= lo < info["n_on"] <= hi
        if ref.lines_num > 500:
            seen_toomuch = True
            assert len(res[t][0]) == 0 and res[t][1].shape == (0, 10)
        elif ref.lines_num:
            assert np.array_equal(det.last_raw[t], np.asarray(ref.linesp_ext).reshape(-1, 4)), t
    assert seen_overflow  # (> 500 raw lines is covered by the golden case synth_256x160_n6_fixed3_dense)


def test_sliding_window_class_golden():
    from metdetpy_b200.detector import SlidingWindow
    g = np.load(os.path.join(GOLDEN, "sliding_window.npz"))
    sw = SlidingWindow(int(g["n"]), g["xs"].shape[1:], np.uint8, force_int=True, calc_std=True)
    assert np.array_equal(sw.sliding_window, np.zeros_like(g["ring"][0]))
    for t, x in enumerate(g["xs"]):
        sw.update(x)
        assert sw.length == g["length"][t] and sw.timer == t + 1
        assert np.array_equal(sw.mean, g["mean"][t]) and sw.mean.dtype == np.uint8
        assert np.array_equal(sw.max, g["max"][t])
        assert np.array_equal(sw.sum, g["sum"][t])
        assert np.array_equal(sw.sliding_window, g["ring"][t])  # utils.py:263-265, the reference's slot order
        assert sw.std == g["std"][t]                            # utils.py:309-321, integer branch
    with pytest.raises(NotImplementedError):
        SlidingWindow(3, (4, 4), float)
    with pytest.raises(AssertionError):
        SlidingWindow(3, (4, 4), np.uint8).std


@pytest.mark.parametrize("n,apply_mask,W,H", [(2, False, 64, 40), (5, True, 96, 64), (12, False, 128, 48), (30, True, 160, 96)])
def test_per_frame_resident_window_state_against_oracle(n, apply_mask, W, H):
    """update()/detect() on the O(1) path (csrc/perframe_kernel.cuh: running sum, prefix max and suffix planes resident
    in HBM) against the CPU oracle, with detect() skipped on some frames, a batched call in the middle (state rebuilt from
    the ring) and the slow path (per_frame_fast = 0: window re-read) beside it."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    rng = np.random.default_rng(100 + n)
    T = 4 * n + 9
    base = rng.integers(20, 60, (H, W))
    frames = np.clip(base[None] + rng.normal(0, 2.5, (T, H, W)), 0, 255).astype(np.uint8)
    for t in range(T):
        x, y = (7 * t) % (W - 14), (3 * t) % (H - 3)
        frames[t, y:y + 2, x:x + 12] = 210
        frames[t, H // 3:H // 3 + 3, W // 4:W // 4 + 3] = 255 if t % 3 else 40
    mask = np.ones((H, W), np.uint8)
    mask[: H // 6, : W // 3] = 0
    kw = dict(adaptive=True, init_value=7, sensitivity="high", area=0.2, interval=1, hough=(6, 6, 4), dy_mask=True)
    ref = O.M3DetectorOracle(n / 10 + 1e-9, 10, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "high", 0.2, 1), HoughLineCfg(6, 6, 4), DynamicCfg(True, 5))
    fast = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=7, apply_mask=apply_mask)
    slow = M3Detector(n / 10 + 1e-9, 10, mask, 10, cfg, None, max_batch=7, apply_mask=apply_mask)
    slow._eng.set_option("per_frame_fast", 0)
    feed = frames if apply_mask else frames * mask[None]
    t = 0
    used_fast = False
    while t < T:
        if t == 2 * n + 3:  # a batched call in the middle: the resident state has to be rebuilt afterwards
            for d in (fast, slow):
                res = d.detect_many(feed[t:t + 7])
            for i in range(7):
                ref.update(frames[t + i] * mask); rl, rc = ref.detect()
            assert fast.last_infos[-1]["bi_threshold"] == ref.bi_threshold
            t += 7
            continue
        ref.update(frames[t] * mask)
        fast.update(feed[t]); slow.update(feed[t])
        if t % 5 != 3:  # every fifth frame is pushed without a detect()
            rl, rc = ref.detect()
            lf, cf = fast.detect()
            used_fast |= int(fast._eng.info("temporal_generation")) == 4
            ls, cs = slow.detect()
            assert int(slow._eng.info("temporal_generation")) != 4
            assert fast.bi_threshold == slow.bi_threshold == ref.bi_threshold, t
            assert np.array_equal(fast.dst, ref.dst), (t, int(np.count_nonzero(fast.dst != ref.dst)))
            assert np.array_equal(slow.dst, ref.dst), t
            raw = np.asarray(ref.linesp_ext).reshape(-1, 4)
            assert np.array_equal(np.asarray(fast.linesp_ext).reshape(-1, 4), raw), t
            assert_nms_equivalent(lf, cf[:, -1], rl, rc[:, -1], raw, t)
        if t in (n - 1, n, 2 * n + 1, T - 1):
            assert np.array_equal(fast.stack.max, ref.stack.max) and np.array_equal(fast.stack.mean, ref.stack.mean), t
            assert np.array_equal(fast.stack.sliding_window, ref.stack.sliding_window), t
        t += 1
    assert used_fast
    fast.close(); slow.close()


def test_per_frame_4k_device_frames_equal_the_batched_path():
    """4K, n = 30: update(on-device frame) + detect() frame by frame == one detect_many over the same frames."""
    import torch
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
    from metdetpy_b200.detector import M3Detector
    W, H, n, T = 3840, 2160, 30, 75
    dev = torch.device("cuda", 0)
    fr = synth.make_stream_device(T, W, H, 30.0, dev, t0=0, loop=4 * T)
    torch.cuda.synchronize()
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(True, 5))
    mask = np.ones((H, W), np.uint8)
    one = M3Detector(n / 30.0 + 1e-9, 30.0, mask, 10, cfg, None)
    many = M3Detector(n / 30.0 + 1e-9, 30.0, mask, 10, cfg, None, max_batch=T)
    res, dst = many.detect_many((fr.data_ptr(), T), on_device=True, return_dst=True)
    lib = one._eng.lib
    from metdetpy_b200._lib import check
    for t in range(T):
        check(lib.mdb_update(one._eng.handle, fr[t].data_ptr(), 1), "update")
        one._timer += 1
        lines, cls = one.detect()
        assert int(one._eng.info("temporal_generation")) == 4
        assert one.bi_threshold == many.last_infos[t]["bi_threshold"], t
        assert np.array_equal(np.asarray(one.linesp_ext).reshape(-1, 4), many.last_raw[t].reshape(-1, 4)), t
        if t % 6 == 0 or t >= T - 2:
            assert np.array_equal(one.dst, dst[t]), t
    one.close(); many.close()


def test_window_longer_than_255_frames_against_oracle():
    """The reference has no limit on int(window_sec * fps) (Detector.py:197): n = 300 runs the generic kernel."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
    from metdetpy_b200.detector import M3Detector
    from oracle import m3_oracle as O
    W, H, n, T, B = 64, 48, 300, 340, 64
    frames = synth.make_stream(T, W, H, 30.0, speed_scale=2.0, thickness=2)
    mask = np.ones((H, W), np.uint8)
    kw = dict(adaptive=True, init_value=7, sensitivity="normal", area=0.2, interval=2, hough=(6, 6, 4), dy_mask=True)
    ref = O.M3DetectorOracle(n / 30.0 + 1e-9, 30.0, mask, 10, backend="numpy", **kw)
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 2), HoughLineCfg(6, 6, 4), DynamicCfg(True, 5))
    det = M3Detector(n / 30.0 + 1e-9, 30.0, mask, 10, cfg, None, max_batch=B)
    assert det.stack_maxsize == n
    for s0 in range(0, T, B):
        res, dst = det.detect_many(frames[s0:s0 + B], return_dst=True)
        for i in range(len(res)):
            t = s0 + i
            ref.update(frames[t]); rl, rc = ref.detect()
            info = det.last_infos[i]
            assert info["bi_threshold"] == ref.bi_threshold, t
            assert info["snr"] == pytest.approx(float(ref.stack.snr), rel=1e-12, abs=0), t
            assert np.array_equal(dst[i], ref.dst), (t, int(np.count_nonzero(dst[i] != ref.dst)))
            raw = np.asarray(ref.linesp_ext).reshape(-1, 4)
            assert_nms_equivalent(res[i][0], res[i][1][:, -1], rl, rc[:, -1], raw, t)
    assert np.array_equal(det.stack.max, ref.stack.max) and np.array_equal(det.stack.mean, ref.stack.mean)
    assert np.array_equal(det.stack.sliding_window, ref.stack.sliding_window)
    det.close()


def test_max_stack_and_merge():
    from metdetpy_b200 import stacker
    rng = np.random.default_rng(2)
    for shape in [(5, 37, 53, 3), (1, 16, 16), (70, 64, 48, 3), (9, 7, 5)]:
        fr = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(stacker.merge_max(fr), fr.max(0))
        box = stacker.MaxImgContainer(chunk=4)
        for f in fr:
            box.append(f)
        assert np.array_equal(box.export(), fr.max(0))

    class Loader:  # the loader protocol _batch_stacker drives (stacker.py:146-175)
        def __init__(self, fr): self.fr, self.i, self.iterations, self.stopped = fr, 0, len(fr), False
        def reset(self, start_frame=None, end_frame=None): self.i, self.iterations = start_frame or 0, (end_frame or len(self.fr)) - (start_frame or 0)
        def start(self): pass
        def stop(self): self.stopped = True
        def pop(self):
            f = self.fr[self.i]; self.i += 1; return f
    fr = rng.integers(0, 256, (20, 30, 40, 3), dtype=np.uint8)
    ld = Loader(fr)
    assert np.array_equal(stacker.max_stacker(ld, 3, 17), fr[3:17].max(0)) and ld.stopped
    assert stacker.max_stacker(Loader(fr[:0])) is None
    assert stacker.MaxImgContainer().export() is None


def test_error_behaviour():
    from metdetpy_b200 import BinaryCfg
    from metdetpy_b200.detector import M3Detector
    det = M3Detector(1, 5, np.ones((32, 48), np.uint8), 10, BinaryCfg(), None, max_batch=4)
    with pytest.raises(ValueError):
        det.update(np.zeros((32, 47), np.uint8))
    with pytest.raises(ValueError):
        det.update(np.zeros((32, 48), np.float32))
    with pytest.raises(Exception):
        det.detect()  # nothing pushed yet
    with pytest.raises(ValueError):
        det.detect_many(np.zeros((5, 32, 48), np.uint8))
    assert det.detect_many(np.zeros((0, 32, 48), np.uint8)) == []
    det.update(np.zeros((32, 48), np.uint8))
    lines, cls = det.detect()  # first frame: empty result, reference shapes
    assert len(lines) == 0 and cls.shape == (0, 10)
    with pytest.raises(ValueError):
        M3Detector(200, 30, np.ones((8, 8), np.uint8), 10, BinaryCfg())  # window 6000 > MDB_MAX_WINDOW


@pytest.mark.parametrize("name,world", [("synth_384x216_n12_dyon_mask", 3), ("clip_192x144_n25", 2),
                                        ("synth_203x157_n3_high", 4)])
def test_time_sharded_equals_reference_golden(name, world):
    """Time-sharding with `world` virtual ranks on one GPU: per-shard noise sums
    (mdb_noise_sums) -> merged -> threshold replay -> per-shard run with halo (mdb_seek,
    mdb_submit_batch_thr).  Every frame must equal the reference's sequential run."""
    from metdetpy_b200 import sharding as S
    g = load_det_case(name)
    cfg, n, T = _cfg(g["cfg"]), g["n"], len(g["frames"])
    eng = S.CudaEngine(g["mask"], n, g["fps"], cfg, max_batch=16)
    assert tuple(eng.roi) == tuple(g["std_roi"])
    shards = S.plan_shards(T, world, n)
    samples = {}
    for sh in shards:
        for tau, s1, s2 in S.local_samples(eng, g["frames"][sh.halo_start:sh.end], sh, n, g["cfg"]["interval"]):
            samples[tau] = S.sigma_from_sums(s1, s2, min(n, tau), eng.roi_pixels)
    c = g["cfg"]
    thr, thr_f, snr = S.replay_thresholds(samples, T, n, adaptive=c["adaptive"], init_value=c["init_value"],
                                          sensitivity=c["sensitivity"], interval=c["interval"])
    assert np.array_equal(thr, g["bi_threshold"])
    assert np.allclose(snr, g["snr"], rtol=1e-12, atol=0)
    for sh in shards:
        h0 = sh.halo_start
        res, dst = eng.detect_chunk(g["frames"][h0:sh.end], h0, thr[h0:sh.end], thr_f[h0:sh.end], snr[h0:sh.end],
                                    want_dst=True)
        for t in range(sh.start, sh.end):
            assert np.array_equal(dst[t - h0], g["dst"][t]), (sh.rank, t)
            lines, cls = res[t - h0]
            ref = ragged_get(g["nms_lines"], g["nms_offs"], t)
            refc = ragged_get(g["cls_pred"], g["nms_offs"], t)
            raw = ragged_get(g["raw_lines"], g["raw_offs"], t)
            assert_nms_equivalent(lines, np.asarray(cls).reshape(-1, 10)[:, -1], ref, refc[:, -1], raw, t)


@pytest.mark.parametrize("name,world,batch", [("synth_384x216_n12_dyon_mask", 3, 16), ("clip_192x144_n25", 2, 32),
                                              ("synth_320x240_n5_dyoff", 4, 7)])
def test_device_chunk_pipeline_equals_reference_golden(name, world, batch):
    """The multi-GPU path of bench.py with `world` virtual ranks on one GPU, frames resident on the device: noise sums of
    every rank's chunk (mdb_noise_sums_dev), native threshold replay (mdb_replay_thresholds), then mdb_reset + mdb_seek
    + a halo batch + the chunk with three batches in flight (sharding.run_chunk).  Thresholds, on-pixel counts, raw
    Hough segments and NMS lines of every frame must equal the reference's sequential run."""
    import torch
    from metdetpy_b200 import sharding as S
    from metdetpy_b200.detector import M3Detector
    g = load_det_case(name)
    cfg, n, T = _cfg(g["cfg"]), g["n"], len(g["frames"])
    H, W = g["frames"].shape[1:]
    dev_frames = torch.from_numpy(g["frames"]).cuda().contiguous()
    base, fb = dev_frames.data_ptr(), H * W
    det = M3Detector(n / g["fps"] + 1e-9, g["fps"], g["mask"], 10, cfg, None, max_batch=batch)
    shards = S.plan_shards(T, world, n)
    c = g["cfg"]

    def pieces(lo, hi):
        return [S.Segment(base + t * fb, min(batch, hi - t), t, history=t) for t in range(lo, hi, batch)]

    samples = np.concatenate([S.chunk_noise_samples(det, pieces(sh.start, sh.end), n, c["interval"], sh.start, sh.end, fb)
                              for sh in shards])
    roi = det.stack.std_roi
    roi_px = (roi[2] - roi[0]) * (roi[3] - roi[1])
    thr, thr_f, snr = S.replay_thresholds_native(samples, roi_px, n, 0, T, adaptive=c["adaptive"], init_value=c["init_value"],
                                                 sensitivity=c["sensitivity"], interval=c["interval"])
    assert np.array_equal(thr, g["bi_threshold"])
    assert np.allclose(snr, g["snr"], rtol=1e-12, atol=0)
    seen = 0
    for sh in shards:
        segs = ([S.Segment(base + sh.halo_start * fb, sh.start - sh.halo_start, sh.halo_start)] if sh.start > sh.halo_start else [])
        if segs and segs[0].T > batch:  # a halo longer than a batch goes in pieces
            segs = pieces(sh.halo_start, sh.start)
        segs += pieces(sh.start, sh.end)
        got = []

        def on_batch(d, sg):
            for i in range(sg.T):
                info = d.last_infos[i]
                k = int(info["n_raw"])
                got.append((sg.t0 + i, int(info["bi_threshold"]), int(info["n_on"]), int(info["lines_num"]),
                            d._eng.raw[i, :k].copy(), d._eng.lines[i, :int(info["n_lines"])].copy(),
                            d._eng.prob[i, :int(info["n_lines"])].copy()))
        S.run_chunk(det, segs, sh, thr, thr_f, snr, on_batch=on_batch)
        assert [x[0] for x in got] == list(range(sh.start, sh.end))
        for t, bt, non, ln, raw, lines, prob in got:
            assert bt == g["bi_threshold"][t], t
            assert non == int(np.count_nonzero(g["dst"][t])), t
            assert ln == g["lines_num"][t], t
            rraw = ragged_get(g["raw_lines"], g["raw_offs"], t)
            if ln <= 500:
                assert np.array_equal(raw, rraw), t
            ref = ragged_get(g["nms_lines"], g["nms_offs"], t)
            refc = ragged_get(g["cls_pred"], g["nms_offs"], t)
            assert_nms_equivalent(lines, prob, ref, refc[:, -1], rraw, t)
            seen += 1
    assert seen == T
    det.close()


def test_config1_bundled_clip_through_the_cuda_class():
    """BASELINE config 1 (test/20220413Red.mp4 with config/m3det_normal.json and test/mask-east.jpg) through the
    CUDA M3Detector with the reference's own call pattern (MetDetPy.py:197-198): the frames are what detect_video
    hands to the detector (resize, gray, real mask, max-merge of exp_frame = 4 decoded frames; window n = 6), the
    golden trajectory is the live reference's.  Spot values of SURVEY 8(c): threshold 5 -> 4, mask_area 474728 and
    std_roi (185,328,355,631) at 960x540, and the one annotated meteor (test/20220413_annotation.json: 2.4 s .. 4.4 s,
    (433,147)-(374,222) on a 1168x655 canvas) is found inside its time span, on its line."""
    from metdetpy_b200.detector import M3Detector
    g = load_det_case("clip_cfg1_480x270_n6")
    H, W = g["frames"].shape[1:]
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None)
    assert det.stack_maxsize == 6
    hits = []
    for t in range(len(g["frames"])):
        det.update(g["frames"][t])
        lines, cls = det.detect()
        _check_frame(det, g, t, lines, cls, det.dst)
        if len(lines):
            hits.append((t, np.asarray(lines)))
    det.close()
    assert g["bi_threshold"][0] == 5 and set(g["bi_threshold"][1:].tolist()) == {4}
    assert hits and hits[0][0] == 20
    eq_fps = 25 / 4
    for t, lines in hits:  # a window of n merged frames ends at frame t
        assert 2.4 - 0.2 <= t / eq_fps <= 4.4 + 6 / eq_fps, t
    # the annotated segment, scaled to this frame size
    sx, sy = W / 1168, H / 655
    p1, p2 = np.array([433 * sx, 147 * sy]), np.array([374 * sx, 222 * sy])
    d = (p2 - p1) / np.linalg.norm(p2 - p1)
    for t, lines in hits:
        for x1, y1, x2, y2 in lines:
            for q in (np.array([x1, y1], float), np.array([x2, y2], float)):
                off = q - p1
                assert abs(off[0] * d[1] - off[1] * d[0]) < 6, (t, lines)          # distance from the annotated line
                assert -10 <= off @ d <= np.linalg.norm(p2 - p1) + 10, (t, lines)  # within its extent
    g2 = load_det_case("clip_cfg1_960x540_n6_range")
    assert int(g2["mask_area"]) == 474728 and tuple(g2["std_roi"]) == (185, 328, 355, 631)
    det = M3Detector(g2["n"] / g2["fps"] + 1e-9, g2["fps"], g2["mask"], 10, _cfg(g2["cfg"]), None, max_batch=22)
    assert tuple(det.stack.std_roi) == (185, 328, 355, 631) and int(det.mask_area) == 474728
    res, dst = det.detect_many(g2["frames"], return_dst=True)
    for t, (lines, cls) in enumerate(res):
        _check_frame(det, g2, t, lines, cls, dst[t], det.last_infos[t])
    assert np.array_equal(np.asarray(res[12][0]), [[337, 152, 351, 126]])
    det.close()
