"""ClassicDetector (MetLib/Detector.py:245-299).
CPU: the restatement (oracle/classic_oracle.py, both backends) against golden trajectories of the
live reference.  GPU: the CUDA path through the C ABI against the same golden vectors and against the
oracle on seeded inputs: thresholds, the difference mask and the Hough segments are bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ragged_get
from oracle import classic_oracle as CO

CASES = ["synth_320x240", "odd_203x157_mask", "dense_256x160_fixed"]


def _load(name):
    g = np.load(os.path.join(GOLDEN, f"classic_{name}.npz"))
    d = {k: g[k] for k in g.files}
    T, H, W = d["frames"].shape
    d["dst"] = np.unpackbits(d["dst_bits"], axis=1)[:, :H * W].reshape(T, H, W) * np.uint8(255)
    return d


def _kw(g):
    adaptive, init_value, area, interval = g["cfg"]
    return dict(adaptive=bool(adaptive), init_value=int(init_value), sensitivity=str(g["sens"]), area=float(area),
                interval=int(interval), hough=tuple(int(v) for v in g["hough"]))


@pytest.mark.parametrize("backend", ["cv2", "numpy"])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name, backend):
    if backend == "cv2":
        pytest.importorskip("cv2")
    g = _load(name)
    det = CO.ClassicDetectorOracle(1.0, float(g["fps"]), g["mask"], 10, backend=backend, **_kw(g))
    assert det.stack_maxsize == 4
    nlines = 0
    for t in range(len(g["frames"])):
        det.update(g["frames"][t])
        lines, cls = det.detect()
        assert det.bi_threshold == g["bi_threshold"][t], t
        assert det.bi_threshold_float == pytest.approx(g["bi_threshold_float"][t], rel=1e-12), t
        assert float(det.stack.snr) == pytest.approx(g["snr"][t], rel=1e-12, abs=0), t
        ref = ragged_get(g["raw_lines"], g["raw_offs"], t)
        assert np.array_equal(np.asarray(lines, np.int32).reshape(-1, 4), ref), t
        if t < 3:
            assert lines == [] and cls == []  # LineDetector.detect(), Detector.py:222-223
        else:
            assert np.asarray(cls).shape == (len(ref), 10) and np.all(np.asarray(cls).reshape(-1, 10)[:, 0] == 1)
        if t >= 3:
            assert np.array_equal(det.dst, g["dst"][t]), t
        nlines += len(ref)
    assert nlines > 0


def _cfg(g):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    k = _kw(g)
    return BinaryCfg(BinaryCoreCfg(k["adaptive"], k["init_value"], k["sensitivity"], k["area"], k["interval"]),
                     HoughLineCfg(*k["hough"]), DynamicCfg(True, 5))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_per_frame_api_matches_reference_golden(name):
    from metdetpy_b200.detector import ClassicDetector
    g = _load(name)
    det = ClassicDetector(123.0, float(g["fps"]), g["mask"], 10, _cfg(g), None)  # window_sec is ignored
    assert det.stack_maxsize == 4
    for t in range(len(g["frames"])):
        det.update(g["frames"][t])
        lines, cls = det.detect()
        assert det.bi_threshold == g["bi_threshold"][t], t
        assert det.bi_threshold_float == pytest.approx(g["bi_threshold_float"][t], rel=1e-12), t
        assert det.stack.snr == pytest.approx(g["snr"][t], rel=1e-12, abs=0), t
        assert np.array_equal(det.dst, g["dst"][t]), (t, int(np.count_nonzero(det.dst != g["dst"][t])))
        ref = ragged_get(g["raw_lines"], g["raw_offs"], t)
        assert np.array_equal(np.asarray(lines, np.int32).reshape(-1, 4), ref), t
        if t < 3:
            assert lines == [] and cls == []
        else:
            assert cls.shape == (len(ref), 10) and np.all(cls[:, 0] == 1) and np.all(cls[:, 1:] == 0)
    det.close()


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 5, 16])
@pytest.mark.parametrize("name", CASES)
def test_gpu_batched_api_matches_reference_golden(name, batch):
    from metdetpy_b200.detector import ClassicDetector
    g = _load(name)
    det = ClassicDetector(1.0, float(g["fps"]), g["mask"], 10, _cfg(g), None, max_batch=batch)
    T = len(g["frames"])
    for s in range(0, T, batch):
        res, dst = det.detect_many(g["frames"][s:s + batch], return_dst=True)
        for i, (lines, cls) in enumerate(res):
            t = s + i
            assert det.last_infos[i]["bi_threshold"] == g["bi_threshold"][t], t
            assert det.last_infos[i]["snr"] == pytest.approx(g["snr"][t], rel=1e-12, abs=0), t
            assert np.array_equal(dst[i], g["dst"][t]), t
            ref = ragged_get(g["raw_lines"], g["raw_offs"], t)
            assert np.array_equal(np.asarray(lines, np.int32).reshape(-1, 4), ref), t
    det.close()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,apply_mask", [(517, 131, False), (128, 96, True), (1920, 1080, False)])
def test_gpu_seeded_streams_against_oracle(W, H, apply_mask):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
    from metdetpy_b200.detector import ClassicDetector
    T, FPS = (24, 30) if W < 1000 else (12, 30)
    frames = synth.make_stream(T, W, H, FPS, speed_scale=3.0 if W < 1000 else 1.0, thickness=2 if W < 1000 else None)
    mask = np.ones((H, W), np.uint8)
    mask[: H // 6, : W // 4] = 0
    feed = frames if apply_mask else frames * mask[None]
    ref = CO.ClassicDetectorOracle(1.0, FPS, mask, 10, adaptive=True, init_value=7, sensitivity="normal", area=0.2,
                                   interval=1, hough=(8, 8, 4), backend="numpy")
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 1), HoughLineCfg(8, 8, 4), DynamicCfg(False, 5))
    det = ClassicDetector(1.0, FPS, mask, 10, cfg, None, max_batch=7, apply_mask=apply_mask)
    got, dsts = [], []
    for s in range(0, T, 7):
        r, d = det.detect_many(feed[s:s + 7], return_dst=True)
        got += r
        dsts += list(d)
    for t in range(T):
        ref.update(frames[t] * mask)
        rl, rc = ref.detect()
        if t >= 3:
            assert np.array_equal(dsts[t], ref.dst), (t, int(np.count_nonzero(dsts[t] != ref.dst)))
        else:
            assert not dsts[t].any()
        assert np.array_equal(np.asarray(got[t][0], np.int32).reshape(-1, 4), np.asarray(rl, np.int32).reshape(-1, 4)), t


@pytest.mark.gpu
def test_gpu_more_than_512_segments_are_all_returned():
    """The reference returns every HoughLinesP segment without a cap (Detector.py:282-292); the library's batched
    outputs hold 512 rows per frame, the rest comes from a second PPHT pass (mdb_get_raw_lines)."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    from metdetpy_b200.detector import ClassicDetector
    W, H, T, FPS = 640, 360, 7, 30
    rng = np.random.default_rng(5)
    frames = np.full((T, H, W), 20, np.uint8)
    for t in range(T):  # a field of short dashes that jumps from frame to frame: hundreds of separate segments
        ys = rng.integers(2, H - 2, 900)
        xs = rng.integers(2, W - 14, 900)
        for y, x in zip(ys, xs):
            frames[t, y, x:x + 9] = 200
    mask = np.ones((H, W), np.uint8)
    ref = CO.ClassicDetectorOracle(1.0, FPS, mask, 10, adaptive=False, init_value=40, sensitivity="normal", area=0.2,
                                   interval=1, hough=(6, 6, 1), backend="cv2")
    cfg = BinaryCfg(BinaryCoreCfg(False, 40, "normal", 0.2, 1), HoughLineCfg(6, 6, 1), DynamicCfg(False, 5))
    det = ClassicDetector(1.0, FPS, mask, 10, cfg, None, max_batch=4)
    det1 = ClassicDetector(1.0, FPS, mask, 10, cfg, None)
    got = []
    for s in range(0, T, 4):
        got += det.detect_many(frames[s:s + 4])
    most = 0
    for t in range(T):
        ref.update(frames[t]); rl, rc = ref.detect()
        det1.update(frames[t]); l1, c1 = det1.detect()
        rl = np.asarray(rl, np.int32).reshape(-1, 4)
        most = max(most, len(rl))
        assert np.array_equal(np.asarray(got[t][0], np.int32).reshape(-1, 4), rl), (t, len(got[t][0]), len(rl))
        assert np.array_equal(np.asarray(l1, np.int32).reshape(-1, 4), rl), t
        if len(rl):
            assert np.asarray(got[t][1]).shape == (len(rl), 10) and np.asarray(c1).shape == (len(rl), 10)
    assert most > 512, most


@pytest.mark.gpu
def test_gpu_exotic_fps_is_refused():
    from metdetpy_b200 import BinaryCfg
    from metdetpy_b200.detector import ClassicDetector
    bad = next(f for f in np.arange(1.3, 3.0, 0.01) if int((4 / f) * f) != 4)
    with pytest.raises(NotImplementedError):
        ClassicDetector(1.0, float(bad), np.ones((32, 32), np.uint8), 10, BinaryCfg(), None)
