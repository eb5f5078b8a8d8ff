"""Clip-level stacking on the GPU: `MaxImgContainer`, `max_stacker` (MetLib/stacker.py:43-49,
:146-175, :197-213), `MergeFunction.max` (MetLib/utils.py:203-204) and `FastGaussianContainer` /
`FastGaussianParam` (MetLib/stacker.py:52-59, MetLib/utils.py:418-513: streaming sum and sum of squares).
Same names and call signatures; frames of any channel layout (the reference stacks full-resolution
colour frames)."""
from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check


def merge_max(image_stack: Sequence[np.ndarray], device: int = 0) -> np.ndarray:
    """MergeFunction.max: element-wise max over a list/array of equally shaped uint8 frames."""
    arr = np.ascontiguousarray(np.asarray(image_stack), np.uint8)
    if arr.ndim < 2 or len(arr) == 0:
        raise ValueError("need a non-empty stack of frames")
    out = np.empty(arr.shape[1:], np.uint8)
    check(_lib.load().mdb_max_stack(arr.ctypes.data, len(arr), out.nbytes, out.ctypes.data, 0, 0, device),
          "merge_max")
    return out


class MaxImgContainer:
    """Running max of appended frames (MetLib/stacker.py:43-49). Frames are buffered and reduced on
    the device in chunks; `container` / `export()` give the current result."""

    def __init__(self, chunk: int = 32, device: int = 0):
        self._buf: list[np.ndarray] = []
        self._acc: Optional[np.ndarray] = None
        self._chunk = chunk
        self._device = device

    def append(self, new_frame: np.ndarray) -> None:
        if self._acc is not None and new_frame.shape != self._acc.shape:
            raise ValueError(f"Expect new frame has the same shape as the base frame "
                             f"{self._acc.shape}, got {new_frame.shape}.")
        if self._buf and new_frame.shape != self._buf[0].shape:
            raise ValueError(f"Expect new frame has the same shape as the base frame "
                             f"{self._buf[0].shape}, got {new_frame.shape}.")
        self._buf.append(np.ascontiguousarray(new_frame, np.uint8))
        if len(self._buf) >= self._chunk:
            self._flush()

    def _flush(self):
        if not self._buf:
            return
        frames = self._buf if self._acc is None else [self._acc] + self._buf
        self._acc = merge_max(frames, self._device)
        self._buf = []

    @property
    def container(self) -> Optional[np.ndarray]:
        self._flush()
        return self._acc

    def export(self):
        return self.container


class FastGaussianParam:
    """Result view of the device accumulation with the reference's attribute names and formulas
    (MetLib/utils.py:418-513): `sum_mu` (uint16), `square_sum` (uint32), `n` (int16), `mu`, `var`.
    The integer fields wrap exactly as the reference's numpy adds do."""

    def __init__(self, sum_mu: np.ndarray, square_sum: np.ndarray, n: np.ndarray, ddof: int = 1):
        self.sum_mu, self.square_sum, self.n, self.ddof = sum_mu, square_sum, n, ddof

    @property
    def mu(self) -> np.ndarray:  # utils.py:454-456
        return np.round(self.sum_mu / self.n)

    @property
    def var(self) -> np.ndarray:  # utils.py:458-465
        sum_mu = np.array(self.sum_mu, dtype=self.square_sum.dtype)
        return (self.square_sum - np.square(sum_mu) / self.n) / (self.n - self.ddof)

    @property
    def shape(self):
        return self.sum_mu.shape


class FastGaussianContainer:
    """FastGaussianContainer (MetLib/stacker.py:52-59): `append(frame)` adds the frame to the running
    per-element sum / sum of squares.  Frames are buffered and reduced on the device in chunks."""

    def __init__(self, chunk: int = 32, device: int = 0):
        self._buf: list[np.ndarray] = []
        self._sum: Optional[np.ndarray] = None
        self._sq: Optional[np.ndarray] = None
        self._count = 0
        self._chunk = chunk
        self._device = device

    def append(self, new_frame: np.ndarray) -> None:
        if self._buf and new_frame.shape != self._buf[0].shape or \
                self._sum is not None and new_frame.shape != self._sum.shape:
            raise ValueError("Expect new frame has the same shape as the base frame")
        self._buf.append(np.ascontiguousarray(new_frame, np.uint8))
        if len(self._buf) >= self._chunk:
            self._flush()

    def _flush(self):
        if not self._buf:
            return
        arr = np.ascontiguousarray(np.stack(self._buf))
        acc = self._sum is not None
        if not acc:
            self._sum = np.empty(arr.shape[1:], np.uint16)
            self._sq = np.empty(arr.shape[1:], np.uint32)
        check(_lib.load().mdb_gauss_stack(arr.ctypes.data, len(arr), arr[0].nbytes, self._sum.ctypes.data,
                                          self._sq.ctypes.data, 0, 0, int(acc), self._device), "FastGaussianContainer")
        self._count += len(arr)
        self._buf = []

    @property
    def container(self) -> Optional[FastGaussianParam]:
        self._flush()
        if self._sum is None:
            return None
        # n: np.ones_like(sum_mu, dtype=int16) added once per frame (utils.py:450-451, :490-493): wraps at 2^15
        n = np.full(self._sum.shape, np.array(self._count).astype(np.int64).astype(np.int16), np.int16)
        return FastGaussianParam(self._sum, self._sq, n)

    def export(self):
        return self.container


def max_stacker(video_loader: Any, start_frame: Optional[int] = None, end_frame: Optional[int] = None,
                logger: Any = None) -> Optional[np.ndarray]:
    """max_stacker (MetLib/stacker.py:197-213) via _batch_stacker's loader protocol (:146-175):
    reset(start,end) / start() / iterations / pop() / stop(). Returns None when no frame came."""
    box = MaxImgContainer()
    try:
        if start_frame is not None or end_frame is not None:
            video_loader.reset(start_frame=start_frame, end_frame=end_frame)
        base_shape = None
        video_loader.start()
        for _ in range(video_loader.iterations):
            img = video_loader.pop()
            if img is None:
                break
            if base_shape is None:
                base_shape = img.shape
            elif base_shape != img.shape:
                raise ValueError(f"Expect new frame has the same shape as the base frame "
                                 f"{base_shape}, got {img.shape}.")
            box.append(img)
    except Exception as e:  # the reference logs and returns what it has (stacker.py:169-171)
        if logger is not None:
            logger.error(e.__repr__())
        return box.container
    finally:
        video_loader.stop()
    return box.container


SUPPORT_BG_ALGO = ["median", "med-of-med", "sigma-clipping", "mean"]  # MetLib/stacker.py:13
_BG_ON_DEVICE = {"mean": 0, "sigma-clipping": 1}
EULER_CONSTANT = 0.5772  # MetLib/utils.py:24


def get_gumbel_mean(n: int) -> float:
    """MetLib/stacker.py:118-126."""
    sqrt2logn = np.sqrt(2 * np.log(n))
    return (sqrt2logn - (np.log(np.log(n)) + np.log(4 * np.pi)) / (2 * sqrt2logn) + EULER_CONSTANT / sqrt2logn)


class MfnrMixContainer:
    """What mfnr_mix_stacker builds with MaxImgContainer + AllImgContainer + FastGaussianContainer
    (MetLib/stacker.py:316-319), kept on the device: `append(frame)` per loader frame, `export(denoise_cfg)` for the
    mixed image.  Frames are shipped in chunks; for sigma clipping they stay resident in HBM for the second pass."""

    def __init__(self, keep_frames: bool, chunk: int = 16, device: int = 0):
        self._buf: list[np.ndarray] = []
        self._h = None
        self._keep, self._chunk, self._device = bool(keep_frames), chunk, device
        self.shape = None
        self.count = 0
        self.stats = None

    def append(self, new_frame: np.ndarray) -> None:
        f = np.ascontiguousarray(new_frame)
        if f.dtype != np.uint8:
            raise ValueError(f"uint8 frames expected, got {f.dtype}")
        if self.shape is None:
            if f.ndim not in (2, 3):
                raise ValueError("frames must be (H, W) or (H, W, C)")
            self.shape = f.shape
            import ctypes as C
            h = C.c_void_p()
            ch = 1 if f.ndim == 2 else f.shape[2]
            check(_lib.load().mdb_mfnr_create(f.shape[0], f.shape[1], ch, int(self._keep), self._device, C.byref(h)),
                  "mfnr create")
            self._h = h
        elif f.shape != self.shape:
            raise ValueError(f"Expect new frame has the same shape as the base frame {self.shape}, got {f.shape}.")
        self._buf.append(f)
        if len(self._buf) >= self._chunk:
            self._flush()

    def _flush(self):
        if self._buf:
            arr = np.ascontiguousarray(np.stack(self._buf))
            check(_lib.load().mdb_mfnr_append(self._h, arr.ctypes.data, len(arr), 0), "mfnr append")
            self.count += len(arr)
            self._buf = []

    def export(self, highlight_preserve: float, blur_ksize: int, bg_algorithm: str, bg_fix_factor: float,
               sigma_high: float = 3.0, sigma_low: float = 3.0) -> np.ndarray:
        import ctypes as C
        self._flush()
        prm = _lib.MfnrParams()
        prm.highlight_preserve, prm.blur_ksize, prm.blur_sigma = float(highlight_preserve), int(blur_ksize), 3.0
        prm.bg_algorithm = _BG_ON_DEVICE[bg_algorithm]
        prm.sigma_high, prm.sigma_low, prm.bg_fix_factor = float(sigma_high), float(sigma_low), float(bg_fix_factor)
        with np.errstate(all="ignore"):
            prm.gumbel_mean = float(get_gumbel_mean(self.count))
        out = np.empty(self.shape, np.uint8)
        st = (C.c_double * 4)()
        check(_lib.load().mdb_mfnr_finish(self._h, C.byref(prm), out.ctypes.data, 0, st), "mfnr finish")
        self.stats = dict(est_bg_var=st[0], gumbel=st[1], highlight_avg_diff=st[2], positive_diffs=int(st[3]))
        return out

    def close(self):
        if self._h is not None:
            _lib.load().mdb_mfnr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mfnr_mix_stacker(video_loader: Any, denoise_cfg: Any, start_frame: Optional[int] = None,
                     end_frame: Optional[int] = None, logger: Any = None) -> Optional[np.ndarray]:
    """mfnr_mix_stacker (MetLib/stacker.py:296-403) on the device, same signature; `denoise_cfg` is the reference's
    DenoiseOption (or anything with its fields).  Built: connect_lines.switch == False with bg_algorithm "mean" or
    "sigma-clipping".  Refused (NotImplementedError): "median" / "med-of-med" (:343-349) and connect_highlight_area
    (:239-294: Lab round trips, Otsu, circular-kernel morphology, contour filling)."""
    algo = denoise_cfg.mfnr_param.bg_algorithm
    assert algo in SUPPORT_BG_ALGO, f"unsupported bg algo! select from {SUPPORT_BG_ALGO}, but {algo} got."
    if algo not in _BG_ON_DEVICE:
        raise NotImplementedError(f"bg_algorithm {algo!r} (MetLib/stacker.py:343-349) is not built on the device")
    if denoise_cfg.connect_lines.switch:
        raise NotImplementedError("connect_lines (connect_highlight_area, MetLib/stacker.py:239-294) is not built on the device")
    box = MfnrMixContainer(keep_frames=algo == "sigma-clipping")
    try:
        try:
            if start_frame is not None or end_frame is not None:
                video_loader.reset(start_frame=start_frame, end_frame=end_frame)
            video_loader.start()
            for _ in range(video_loader.iterations):
                img = video_loader.pop()
                if img is None:
                    break
                box.append(img)
        except Exception as e:  # the reference logs and goes on with what it has (stacker.py:169-171)
            if logger is not None:
                logger.error(e.__repr__())
        finally:
            video_loader.stop()
        box._flush()
        if box.count == 0:
            return None
        # stacker.py:333-336: single_sigma_clipping is always called with sigma 3.0 / 3.0
        out = box.export(denoise_cfg.highlight_preserve, denoise_cfg.blur_ksize, algo, denoise_cfg.mfnr_param.bg_fix_factor)
        if logger is not None and hasattr(logger, "debug"):
            logger.debug(f"highlight fix factor = {box.stats['est_bg_var'] * box.stats['gumbel'] * denoise_cfg.mfnr_param.bg_fix_factor:.4f}")
        return out
    finally:
        box.close()
