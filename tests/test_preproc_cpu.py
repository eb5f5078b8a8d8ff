"""Loader preprocessing, CPU side: the numpy restatement against the
golden vectors of the live reference's Transform / MergeFunction (tests/golden/preproc.npz), against
cv2 itself on seeded inputs, and the library's host-side tap tables against the restatement."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import preproc_oracle as P

cv2 = pytest.importorskip("cv2")


def _golden():
    return np.load(os.path.join(GOLDEN, "preproc.npz"))


def test_oracle_reproduces_reference_transform_golden():
    g = _golden()
    assert len(g["names"]) >= 4
    for name in g["names"]:
        W, H, gray, exp = (int(v) for v in g[f"{name}_cfg"])
        out = P.preprocess_stream(g[f"{name}_frames"], (W, H), bool(gray), g[f"{name}_mask"], exp)
        assert out.dtype == np.uint8 and np.array_equal(out, g[f"{name}_out"]), name


@pytest.mark.parametrize("W0,H0,W,H,C", [(384, 216, 96, 54, 3), (200, 120, 97, 61, 3), (64, 48, 160, 90, 3),
                                         (333, 211, 100, 80, 1), (96, 54, 96, 108, 3), (40, 40, 1, 1, 1),
                                         (3, 2, 64, 48, 3)])
def test_resize_and_gray_restatement_equals_cv2(W0, H0, W, H, C):
    """cv2 4.13 semantics the restatement follows (bit-exact): 8-bit INTER_LINEAR resize, BGR2GRAY."""
    rng = np.random.default_rng(W0 * 7 + W)
    img = rng.integers(0, 256, (H0, W0, C), dtype=np.uint8)
    if C == 1:
        img = img[..., 0]
    assert np.array_equal(P.resize_linear_u8(img, (W, H)), cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR))
    if C == 3:
        assert np.array_equal(P.bgr2gray_u8(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_gray_restatement_exhaustive_grid():
    v = np.arange(256, dtype=np.uint8)
    B, G, R = np.meshgrid(v[::3], v[::5], v[::7], indexing="ij")
    img = np.ascontiguousarray(np.stack([B, G, R], -1).reshape(-1, 1, 3))
    assert np.array_equal(P.bgr2gray_u8(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_library_tap_tables_equal_restatement():
    """mdb_preproc_axis_taps is host-only code of the CUDA library: loadable and callable without a GPU."""
    from metdetpy_b200 import _lib
    lib = _lib.load()
    for dst, src in [(960, 3840), (540, 2160), (1000, 3840), (563, 2160), (960, 640), (97, 200), (1, 40), (64, 3)]:
        for clamp in (0, 1):
            arrs = [np.zeros(dst, np.int32) for _ in range(4)]
            assert lib.mdb_preproc_axis_taps(dst, src, clamp, *[a.ctypes.data for a in arrs]) == 0
            ref = P.axis_taps(dst, src, bool(clamp))
            for a, b in zip(arrs, ref):
                assert np.array_equal(a, b), (dst, src, clamp)


def test_transform_plan_validation_needs_no_gpu():
    from metdetpy_b200.imgproc import Transform
    tr = Transform()
    tr.opencv_BGR2GRAY()
    tr.opencv_resize([8, 8])
    with pytest.raises(NotImplementedError):
        tr._plan((16, 16, 3))
    tr = Transform()
    tr.opencv_resize([8, 4])
    with pytest.raises(NotImplementedError):  # colour output is not part of the detector path
        tr._plan((16, 16, 3))
    tr.opencv_BGR2GRAY()
    tr.mask_with(np.ones((4, 8), np.uint8))
    assert tr._plan((16, 16, 3)) [:3] == (3, False, (8, 4))
    with pytest.raises(NotImplementedError):
        Transform().opencv_resize([8, 8], resize_interpolation=3)
