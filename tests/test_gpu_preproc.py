"""Loader preprocessing on the device through the C ABI: bit-exact
against the golden vectors of the live reference's Transform, against the CPU restatement, and against
cv2 at the sizes the reference runs (4K / 1080p source -> 960x540)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import preproc_oracle as P

pytestmark = pytest.mark.gpu


def _transform(W0, H0, W, H, gray, mask, rgb=False):
    from metdetpy_b200.imgproc import Transform
    tr = Transform()
    if (W0, H0) != (W, H):
        tr.opencv_resize([W, H])
    if gray:
        tr.opencv_RGB2GRAY() if rgb else tr.opencv_BGR2GRAY()
    if mask is not None:
        tr.mask_with(mask)
    return tr


def test_reference_transform_golden():
    g = np.load(os.path.join(GOLDEN, "preproc.npz"))
    for name in g["names"]:
        W, H, gray, exp = (int(v) for v in g[f"{name}_cfg"])
        frames, mask = g[f"{name}_frames"], g[f"{name}_mask"]
        tr = _transform(frames.shape[2], frames.shape[1], W, H, bool(gray), mask)
        out = tr.exec_transform_many(frames, exp)
        assert out.dtype == np.uint8 and np.array_equal(out, g[f"{name}_out"]), name
        one = tr.exec_transform(frames[0]) if exp == 1 else None  # reference per-frame signature
        if one is not None:
            assert np.array_equal(one, g[f"{name}_out"][0]), name
        tr.close()


@pytest.mark.parametrize("W0,H0,W,H,C,exp,T", [(200, 120, 97, 61, 3, 1, 3), (64, 48, 160, 90, 3, 2, 5),
                                               (333, 211, 100, 80, 1, 3, 7), (96, 54, 96, 54, 3, 4, 8),
                                               (3, 2, 64, 48, 3, 1, 2), (640, 480, 960, 540, 3, 1, 2)])
def test_seeded_inputs_against_oracle(W0, H0, W, H, C, exp, T):
    rng = np.random.default_rng(W0 + 13 * W + exp)
    frames = rng.integers(0, 256, (T, H0, W0, C), dtype=np.uint8)
    if C == 1:
        frames = frames[..., 0]
    mask = (rng.random((H, W)) > 0.3).astype(np.uint8)
    for m in (mask, None):
        tr = _transform(W0, H0, W, H, C == 3, m)
        out = tr.exec_transform_many(frames, exp)
        ref = P.preprocess_stream(frames, (W, H), C == 3, m, exp)
        assert np.array_equal(out, ref), (int(np.count_nonzero(out != ref)), m is None)
        tr.close()


def test_rgb_order():
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 40, 60, 3), dtype=np.uint8)
    tr = _transform(60, 40, 30, 20, True, None, rgb=True)
    out = tr.exec_transform_many(frames, 1)
    ref = P.preprocess_stream(frames[..., ::-1], (30, 20), True, None, 1)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("W0,H0", [(3840, 2160), (1920, 1080)])
def test_full_size_against_cv2(W0, H0):
    """The reference's default runtime size (960x540) from 4K / 1080p BGR frames, against cv2 directly."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(W0)
    T, exp = 4, 2
    base = cv2.GaussianBlur(rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8), (0, 0), 2.0)
    frames = np.stack([np.clip(base.astype(np.int16) + rng.integers(-30, 30, base.shape, dtype=np.int16), 0, 255).astype(np.uint8)
                       for _ in range(T)])
    mask = np.ones((540, 960), np.uint8)
    mask[400:, :300] = 0
    tr = _transform(W0, H0, 960, 540, True, mask)
    out = tr.exec_transform_many(frames, exp)
    ref = []
    for s in range(0, T, exp):
        grp = [cv2.cvtColor(cv2.resize(f, (960, 540), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) * mask
               for f in frames[s:s + exp]]
        ref.append(np.max(grp, axis=0))
    assert np.array_equal(out, np.stack(ref))


def test_preprocessed_frames_feed_the_detector_without_leaving_the_gpu():
    """Transform(keep_on_device) -> M3Detector.detect_many(on_device=True) equals host preprocessing
    (oracle) -> detect_many(host frames)."""
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
    from metdetpy_b200.detector import M3Detector
    W0, H0, W, H, T, n, FPS = 768, 432, 384, 216, 24, 5, 30
    gray = synth.make_stream(T, W0, H0, FPS, speed_scale=3.0, thickness=3)
    bgr = np.stack([gray, np.roll(gray, 1, axis=2), gray // 2 + 10], axis=-1).astype(np.uint8)
    mask = np.ones((H, W), np.uint8)
    mask[:20] = 0
    cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(True, 5))
    tr = _transform(W0, H0, W, H, True, mask)
    dev = tr.exec_transform_many(bgr, 1, keep_on_device=True)
    assert dev.shape == (T, H, W)
    det_a = M3Detector(n / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=T)
    res_a, dst_a = det_a.detect_many((dev.ptr, T), on_device=True, return_dst=True)
    host = P.preprocess_stream(bgr, (W, H), True, mask, 1)
    det_b = M3Detector(n / FPS + 1e-9, FPS, mask, 10, cfg, None, max_batch=T)
    res_b, dst_b = det_b.detect_many(host, return_dst=True)
    assert np.array_equal(dst_a, dst_b)
    for (la, ca), (lb, cb) in zip(res_a, res_b):
        assert np.array_equal(np.asarray(la).reshape(-1, 4), np.asarray(lb).reshape(-1, 4))
    assert sum(len(r[0]) for r in res_a) > 0


def test_errors():
    from metdetpy_b200.imgproc import Transform
    tr = Transform()
    tr.opencv_resize([8, 8])
    with pytest.raises(NotImplementedError):
        tr.exec_transform(np.zeros((16, 16, 3), np.uint8))  # colour output unsupported
    tr.opencv_BGR2GRAY()
    with pytest.raises(ValueError):
        tr.exec_transform(np.zeros((16, 16, 3), np.float32))
    tr.mask_with(np.ones((4, 4), np.uint8))
    with pytest.raises(ValueError):
        tr.exec_transform(np.zeros((16, 16, 3), np.uint8))  # mask shape mismatch
