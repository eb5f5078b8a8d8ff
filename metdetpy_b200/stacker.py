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


class _DeviceAccumulator:
    """Frames are shipped in chunks to a device-resident accumulator (max, uint16 sum, uint32 sum of squares per element:
    `mdb_mfnr_append`); nothing but the new frames crosses PCIe, the running result is read back on demand."""

    def __init__(self, chunk: int, device: int):
        self._buf: list[np.ndarray] = []
        self._h = None
        self._chunk, self._device = chunk, device
        self.shape = None
        self.count = 0

    def append(self, new_frame: np.ndarray) -> None:
        f = np.asarray(new_frame)
        if f.dtype != np.uint8:
            raise ValueError(f"uint8 frames expected, got {f.dtype}")
        if self.shape is None:
            import ctypes as C
            if f.ndim not in (2, 3) or (f.ndim == 3 and f.shape[2] > 4):
                raise ValueError("frames must be (H, W) or (H, W, C <= 4)")
            self.shape = f.shape
            h = C.c_void_p()
            check(_lib.load().mdb_mfnr_create(f.shape[0], f.shape[1], 1 if f.ndim == 2 else f.shape[2], 0, self._device,
                                              C.byref(h)), "stack accumulator")
            self._h = h
        elif f.shape != self.shape:
            raise ValueError(f"Expect new frame has the same shape as the base frame {self.shape}, got {f.shape}.")
        self._buf.append(np.ascontiguousarray(f))
        if len(self._buf) >= self._chunk:
            self._flush()

    def _flush(self):
        if self._buf:
            arr = np.ascontiguousarray(np.stack(self._buf))
            check(_lib.load().mdb_mfnr_append(self._h, arr.ctypes.data, len(arr), 0), "stack accumulator")
            self.count += len(arr)
            self._buf = []

    def _read(self, want_max=False, want_sums=False):
        self._flush()
        if self.count == 0:
            return None, None, None
        mx = np.empty(self.shape, np.uint8) if want_max else None
        sm = np.empty(self.shape, np.uint16) if want_sums else None
        sq = np.empty(self.shape, np.uint32) if want_sums else None
        p = lambda a: None if a is None else a.ctypes.data
        check(_lib.load().mdb_mfnr_stats(self._h, p(mx), p(sm), p(sq), None), "stack accumulator")
        return mx, sm, sq

    def close(self):
        if self._h is not None:
            _lib.load().mdb_mfnr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MaxImgContainer(_DeviceAccumulator):
    """Running max of appended frames (MetLib/stacker.py:43-49), accumulated on the device; `container` / `export()`
    give the current result."""

    def __init__(self, chunk: int = 32, device: int = 0):
        super().__init__(chunk, device)

    @property
    def container(self) -> Optional[np.ndarray]:
        return self._read(want_max=True)[0]

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

    # arithmetic of the reference class (utils.py:485-509), numpy dtypes and wrap-around as there; host-side result
    # objects only -- the accumulation itself runs on the device
    def __add__(self, g2: "FastGaussianParam") -> "FastGaussianParam":
        assert isinstance(g2, FastGaussianParam), "unacceptable object"
        assert self.ddof == g2.ddof, "unmatched var calculation!"
        return FastGaussianParam(self.sum_mu + g2.sum_mu, self.square_sum + g2.square_sum, self.n + g2.n, self.ddof)

    def __sub__(self, g2: "FastGaussianParam") -> "FastGaussianParam":
        assert isinstance(g2, FastGaussianParam), "unacceptable object"
        assert self.ddof == g2.ddof, "unmatched var calculation!"
        assert (self.n - g2.n).any() >= 0, "generate n<0 fistribution!"
        return FastGaussianParam(self.sum_mu - g2.sum_mu, self.square_sum - g2.square_sum, self.n - g2.n, self.ddof)

    def mask(self, mask_pos: np.ndarray) -> None:
        assert mask_pos.dtype == np.dtype("bool"), "Invalid mask!"
        self.sum_mu *= mask_pos
        self.square_sum *= mask_pos
        self.n = np.array(mask_pos, dtype=np.uint16)

    def apply_zero_var(self, full_img: "FastGaussianParam") -> None:
        zero_pos = (self.n == 0)
        self.n[zero_pos] = full_img.n[zero_pos]
        self.sum_mu[zero_pos] = full_img.sum_mu[zero_pos]
        self.square_sum[zero_pos] = full_img.square_sum[zero_pos]

    def upscale(self) -> None:
        up = {np.dtype("uint8"): np.dtype("uint16"), np.dtype("uint16"): np.dtype("uint32"), np.dtype("uint32"): np.dtype("uint64")}
        self.sum_mu = np.array(self.sum_mu, dtype=up.get(self.sum_mu.dtype, float))
        self.square_sum = np.array(self.square_sum, dtype=up.get(self.square_sum.dtype, float))


class FastGaussianContainer(_DeviceAccumulator):
    """FastGaussianContainer (MetLib/stacker.py:52-59): `append(frame)` adds the frame to the running
    per-element sum / sum of squares, accumulated on the device."""

    def __init__(self, chunk: int = 32, device: int = 0):
        super().__init__(chunk, device)

    @property
    def container(self) -> Optional[FastGaussianParam]:
        _, sm, sq = self._read(want_sums=True)
        if sm is None:
            return None
        # n: np.ones_like(sum_mu, dtype=int16) added once per frame (utils.py:450-451, :490-493): wraps at 2^15
        n = np.full(sm.shape, np.array(self.count).astype(np.int64).astype(np.int16), np.int16)
        return FastGaussianParam(sm, sq, n)

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
_BG_ON_DEVICE = {"mean": 0, "sigma-clipping": 1, "median": 2, "med-of-med": 3}
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

    def expect(self, frames: int) -> None:
        """How many frames the clip will deliver: the device memory for them is then reserved in one piece."""
        self._expect = max(0, int(frames))

    def _flush(self):
        if self._buf:
            arr = np.ascontiguousarray(np.stack(self._buf))
            if self._keep and self.count == 0 and getattr(self, "_expect", 0):
                check(_lib.load().mdb_mfnr_reserve(self._h, self._expect), "mfnr reserve")
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
        prm.med_block_size = int(self.count ** (1 / 2))  # median_of_medians' default block size (stacker.py:68-69)
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
    DenoiseOption (or anything with its fields).  Built: connect_lines.switch == False with every bg_algorithm ("mean",
    "sigma-clipping", "median", "med-of-med").  Refused (NotImplementedError): connect_highlight_area (:239-294: Lab
    round trips, Otsu, circular-kernel morphology, contour filling)."""
    algo = denoise_cfg.mfnr_param.bg_algorithm
    assert algo in SUPPORT_BG_ALGO, f"unsupported bg algo! select from {SUPPORT_BG_ALGO}, but {algo} got."
    if denoise_cfg.connect_lines.switch:
        raise NotImplementedError("connect_lines (connect_highlight_area, MetLib/stacker.py:239-294) is not built on the device")
    box = MfnrMixContainer(keep_frames=algo != "mean")
    try:
        try:
            if start_frame is not None or end_frame is not None:
                video_loader.reset(start_frame=start_frame, end_frame=end_frame)
            video_loader.start()
            box.expect(int(video_loader.iterations))
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
