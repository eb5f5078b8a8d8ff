"""CPU restatement (numpy, integer arithmetic) of the reference's loader preprocessing -- TEST
INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline may import it; the
product never does).

Path restated: `Transform.exec_transform` (MetLib/imgproc.py:129-139)
with the chain the video loader builds (MetLib/videoloader.py:300-308):
    cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)   imgproc.py:82-85
 -> cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)                    imgproc.py:87-88
 -> img * mask                                               imgproc.py:96-101
and the exposure merge `MergeFunction.max` = np.max(frames, axis=0) (MetLib/utils.py:203-204,
videoloader.py:388) over `exp_frame` consecutive frames.

Third-party arithmetic: opencv-python (requirements.txt:1, >= 4.9.0; 4.13.0 here).  Its 8-bit
INTER_LINEAR resize and BGR2GRAY are fixed-point; the formulas below were checked bit-exact against
cv2 4.13.0 (IPP on and off give identical results) for down- and up-scaling, integer and
non-integer ratios, 1 and 3 channels (tests/test_preproc_cpu.py keeps doing that), and are pinned
on golden vectors produced by the live reference's Transform (tests/golden/preproc.npz).

  resize, per axis:  f = (float32)((d + 0.5) * (double)src/dst - 0.5);  s = floor(f);  f -= s
      x axis only:   s < 0 -> (s, f) = (0, 0);   s >= src-1 -> (s, f) = (src-1, 0)
      y axis:        the two source rows are clip(s, 0, src-1) and clip(s+1, 0, src-1), f is kept
      weights:       w1 = cvRound(f * 2048), w0 = cvRound((1 - f) * 2048)   (float32, half to even)
      horizontal:    h = S[s] * a0 + S[s+1] * a1                            (int32)
      vertical:      out = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2
  gray:              (B*3735 + G*19235 + R*9798 + 16384) >> 15
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS
BY15, GY15, RY15, GRAY_SHIFT = 3735, 19235, 9798, 15


def axis_taps(dst: int, src: int, clamp_fraction: bool):
    """Source index pair and 11-bit weights of every destination coordinate (see module docstring).
    Returns int32 arrays (s0, s1, w0, w1)."""
    scale = float(src) / float(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    s0 = np.clip(s, 0, src - 1).astype(np.int32)
    s1 = np.clip(s + 1, 0, src - 1).astype(np.int32)
    return s0, s1, w0, w1


def resize_linear_u8(img: np.ndarray, dsize) -> np.ndarray:
    """cv2.resize(img, dsize=(W, H), interpolation=cv2.INTER_LINEAR) for uint8, 1 or 3 channels."""
    W, H = int(dsize[0]), int(dsize[1])
    H0, W0 = img.shape[:2]
    if (W0, H0) == (W, H):
        return img.copy()
    sx0, sx1, a0, a1 = axis_taps(W, W0, True)
    sy0, sy1, b0, b1 = axis_taps(H, H0, False)
    im = img.astype(np.int64)
    if im.ndim == 2:
        im = im[..., None]
    h = im[:, sx0, :] * a0[None, :, None] + im[:, sx1, :] * a1[None, :, None]
    h0, h1 = h[sy0], h[sy1]
    out = (((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out if img.ndim == 3 else out[..., 0]


def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8."""
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    return ((b * BY15 + g * GY15 + r * RY15 + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT).astype(np.uint8)


def preprocess_frame(img: np.ndarray, dsize, grayscale: bool, mask) -> np.ndarray:
    """Transform.exec_transform with resize -> BGR2GRAY -> mask_with (each step optional as in
    videoloader.py:300-308)."""
    H0, W0 = img.shape[:2]
    if (W0, H0) != (int(dsize[0]), int(dsize[1])):
        img = resize_linear_u8(img, dsize)
    if grayscale and img.ndim == 3:
        img = bgr2gray_u8(img)
    if mask is not None:
        img = img * mask
    return img


def preprocess_stream(frames, dsize, grayscale: bool, mask, exp_frame: int) -> np.ndarray:
    """What the loader hands to the detector: groups of `exp_frame` consecutive preprocessed frames
    merged by MergeFunction.max (a shorter last group is merged as it is, videoloader.py:380-388)."""
    out = []
    for s in range(0, len(frames), exp_frame):
        group = [preprocess_frame(f, dsize, grayscale, mask) for f in frames[s:s + exp_frame]]
        out.append(group[0] if len(group) == 1 else np.max(group, axis=0, keepdims=True)[0])
    return np.stack(out)
