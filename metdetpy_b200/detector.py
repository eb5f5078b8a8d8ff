"""Host-side mirror of the reference's detector interface for the M3 line-detector path.

Same class names, constructor arguments, methods, attributes and return types as
MetLib/Detector.py (`BaseDetector` :130-157, `LineDetector` :160-242, `M3Detector` :302-448,
`SNR_SW` :34-127) and MetLib/utils.py (`SlidingWindow` :225-321, `EMA` :324-368,
`lineset_nms` :780-839), so `MetDetPy.detect_video` (MetDetPy.py:137-142, :197-198) can drive it
unchanged.  All pixel work happens in libmetdet_b200.so (CUDA, sm_100a) through the C ABI in
include/metdet_b200.h; this module only marshals buffers.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
from abc import ABCMeta, abstractmethod
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MAX_LINES, Config, FrameInfo, check
from .config import BinaryCfg

NUM_LINES_TOOMUCH = 500  # MetLib/Detector.py:30
DEFAULT_INIT_VALUE = 5  # MetLib/Detector.py:31
PI = np.pi / 180.0  # MetLib/utils.py:22
_SENS_CODE = {"low": 0, "normal": 1, "high": 2}
_INFO_DTYPE = np.dtype([("timer", "<i8"), ("bi_threshold", "<i4"), ("n_on", "<i4"), ("bi_threshold_float", "<f8"),
                        ("snr", "<f8"), ("dst_sum", "<f8"), ("gap", "<f8"), ("lines_num", "<i4"), ("n_raw", "<i4"),
                        ("n_lines", "<i4"), ("len_ties", "<i4")])


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def select_subarea(mask: np.ndarray, area: float) -> tuple[int, int, int, int]:
    """SNR_SW.select_subarea (MetLib/Detector.py:93-122): centre ROI of relative area `area`, moved
    up in 10-row steps while the un-masked share does not drop. Returns std_roi (r0, c0, r1, c1)."""
    h, w = mask.shape[:2]
    if area == 0:
        raise ValueError("binary.area == 0 is not supported (the reference's own path is broken "
                         "for it, MetLib/Detector.py:105-107)")
    rate = area ** (1 / 2)
    sub_h, sub_w = int(h * rate), int(w * rate)
    if sub_h < 1 or sub_w < 1:
        raise ValueError(f"binary.area={area} gives an empty noise ROI for a {w}x{h} frame")
    r0, c0 = (h - sub_h) // 2, (w - sub_w) // 2
    tot = sub_h * sub_w
    ratio = np.sum(mask[r0:r0 + sub_h, c0:c0 + sub_w]) / tot
    while ratio < 1:
        r0 -= 10
        new_ratio = np.sum(mask[r0:r0 + sub_h, c0:c0 + sub_w]) / tot
        if new_ratio < ratio or r0 < 0:
            r0 += 10
            break
        ratio = new_ratio
    return (r0, c0, r0 + sub_h, c0 + sub_w)


def lineset_nms(lines: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """lineset_nms (MetLib/utils.py:780-839): the greedy pass runs in the library's host C++; the visiting order is the
    reference's own `np.argsort(length_sqr)[::-1]` (utils.py:804), evaluated by numpy on this host on the same int32
    values, so that equal lengths are ordered exactly as the reference orders them here."""
    lines = np.ascontiguousarray(lines, np.int32).reshape(-1, 4)
    n = len(lines)
    out = np.empty((max(n, 1), 4), np.int32)
    prob = np.empty(max(n, 1), np.float64)
    k = C.c_int32(0)
    length_sqr = np.power((lines[:, 3] - lines[:, 1]), 2) + np.power((lines[:, 2] - lines[:, 0]), 2)
    order = np.ascontiguousarray(np.argsort(length_sqr)[::-1], np.int32)
    check(_lib.load().mdb_lineset_nms_ordered(_ptr(lines), n, _ptr(order), _ptr(out), _ptr(prob), C.byref(k)), "lineset_nms")
    return out[:k.value].copy(), prob[:k.value].copy()


class _RaggedRows:
    """Per-frame row blocks of one flat (K, 4) array: `rows[i]` -> the (k_i, 4) segments of frame i."""

    def __init__(self, flat: np.ndarray, offs: np.ndarray):
        self.flat, self.offs = flat, offs

    def __len__(self) -> int:
        return len(self.offs) - 1

    def __getitem__(self, i: int) -> np.ndarray:
        if i < 0:
            i += len(self)
        return self.flat[self.offs[i]:self.offs[i + 1]]


class EMA:
    """EMA (MetLib/utils.py:324-368). Scalar host recurrence kept for API compatibility; inside the
    detector the same recurrence runs on the device (threshold_kernel)."""

    def __init__(self, momentum: float = 0.99, warmup_speed: float = 1) -> None:
        assert 0 <= momentum <= 1, "momentum should be [0,1]"
        self.init_momentum = momentum
        self.cur_momentum = momentum
        self.cur_value = 0
        self.t = 0
        self.warmup_speed = warmup_speed

    def update(self, value) -> None:
        if self.warmup_speed:
            self.adjust_weight()
        self.cur_value = self.cur_momentum * self.cur_value + (1 - self.cur_momentum) * value
        self.t += 1

    def adjust_weight(self) -> None:
        k = self.t * (1 - self.init_momentum) * self.warmup_speed
        if k < 1:
            self.cur_momentum = self.init_momentum * (1 - (1 - k) ** 2)
        else:
            self.warmup_speed = 0
            self.cur_momentum = self.init_momentum


class _Engine:
    """Owns one mdb_handle and its output staging buffers."""

    def __init__(self, mask: np.ndarray, n: int, *, adaptive: bool, init_value: int,
                 sensitivity: str, interval: int, roi, hough, dy_mask: bool, max_batch: int,
                 device: int, apply_mask: bool, detector: int = 0):
        self.lib = _lib.load()
        if self.lib.mdb_device_count() < 1:
            raise _lib.MetDetError("no CUDA device visible: metdetpy_b200 has no CPU fallback")
        mask = np.ascontiguousarray(mask, np.uint8)
        if mask.ndim != 2:
            raise ValueError("mask must be a (H, W) uint8 array of {0,1}")
        self.H, self.W = mask.shape
        self.n = n
        self.max_batch = max_batch
        cfg = Config()
        cfg.width, cfg.height, cfg.window = self.W, self.H, n
        cfg.adaptive, cfg.init_value = int(bool(adaptive)), int(init_value)
        if sensitivity not in _SENS_CODE:
            raise KeyError(sensitivity)
        cfg.sensitivity = _SENS_CODE[sensitivity]
        cfg.nz_interval = int(interval)
        for i in range(4):
            cfg.roi[i] = int(roi[i])
        cfg.hough_threshold, cfg.hough_min_len, cfg.hough_max_gap = (int(v) for v in hough)
        cfg.dy_mask = int(bool(dy_mask))
        cfg.max_batch, cfg.device, cfg.apply_mask = int(max_batch), int(device), int(bool(apply_mask))
        cfg.detector = int(detector)
        self.handle = C.c_void_p()
        check(self.lib.mdb_create(C.byref(cfg), _ptr(mask), C.byref(self.handle)), "mdb_create")
        T = max_batch
        self.infos = (FrameInfo * T)()
        self.lines = np.zeros((T, MAX_LINES, 4), np.int32)
        self.prob = np.zeros((T, MAX_LINES), np.float64)
        self.raw = np.zeros((T, MAX_LINES, 4), np.int32)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.mdb_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check_frame(self, frame: np.ndarray) -> np.ndarray:
        if frame.dtype != np.uint8 or frame.shape[-2:] != (self.H, self.W):
            raise ValueError(f"expected uint8 frame(s) of shape (..., {self.H}, {self.W}), "
                             f"got {frame.dtype} {frame.shape}")
        return np.ascontiguousarray(frame)

    def set_option(self, name: str, value: int):
        check(self.lib.mdb_set_option(self.handle, name.encode(), int(value)), "mdb_set_option")

    def info(self, name: str) -> float:
        v = C.c_double()
        check(self.lib.mdb_get_info(self.handle, name.encode(), C.byref(v)), "mdb_get_info")
        return v.value

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        check(self.lib.mdb_get_stream(self.handle, C.byref(p)), "mdb_get_stream")
        return p.value or 0

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.lib.mdb_get_launch_count(self.handle, C.byref(v)), "mdb_get_launch_count")
        return v.value

    def fused_time(self) -> tuple[float, int]:
        ms, nl = C.c_float(), C.c_int32()
        check(self.lib.mdb_get_fused_time(self.handle, C.byref(ms), C.byref(nl)), "mdb_get_fused_time")
        return ms.value, nl.value


class SlidingWindow(object):
    """SlidingWindow (MetLib/utils.py:225-321) for uint8 frames, backed by the device frame ring:
    `update`, `max`, `mean`, `sum`, `length`, `timer`, `cur_index`, `refresh_max`, `sliding_window`, `std`
    (with `calc_std=True`).  Only the reference's `dtype=np.uint8, force_int=True` mode (the one the detectors
    use, Detector.py:53-57, :212-215) exists; anything else raises."""

    def __init__(self, n: int, size: Sequence[int], dtype: type = np.uint8, force_int: bool = True,
                 calc_std: bool = False, device: int = 0) -> None:
        if np.dtype(dtype) != np.uint8 or not force_int:
            # utils.py:248-252: other element types switch the sums to float64; every SlidingWindow the reference
            # constructs (Detector.py:53-57, :67-71, :213-216, :536-539) is dtype=np.uint8, force_int=True
            raise NotImplementedError("device SlidingWindow supports dtype=uint8, force_int=True (the mode of every "
                                      "SlidingWindow the reference constructs); MetLib/utils.py:248-252 float mode is not built")
        if len(size) != 2:
            raise ValueError("size must be (H, W)")
        self.n = n
        self.size = tuple(size)
        self.dtype = dtype
        self.force_int = force_int
        self.calc_std = calc_std
        h, w = self.size
        self._eng = _Engine(np.ones((h, w), np.uint8), n, adaptive=False, init_value=0,
                            sensitivity="normal", interval=0, roi=(0, 0, 1, 1), hough=(1, 1, 1),
                            dy_mask=False, max_batch=1, device=device, apply_mask=False)
        self.timer = 0
        self.cur_index = 0

    def update(self, new_frame: np.ndarray) -> None:
        f = self._eng.check_frame(new_frame)
        check(self._eng.lib.mdb_update(self._eng.handle, _ptr(f), 0), "SlidingWindow.update")
        self.timer += 1
        self.cur_index = (self.timer - 1) % self.n

    def _stack(self, which: int):
        h, w = self.size
        if self.timer == 0:
            return np.zeros((h, w), np.uint32 if which == 2 else np.uint8)
        out = np.empty((h, w), np.uint32 if which == 2 else np.uint8)
        args = [None, None, None]
        args[which] = _ptr(out)
        check(self._eng.lib.mdb_get_stack(self._eng.handle, *args), "SlidingWindow")
        return out

    @property
    def length(self) -> int:
        return min(self.n, self.timer)

    @property
    def max(self) -> np.ndarray:
        return self._stack(0)

    @property
    def mean(self) -> np.ndarray:
        return self._stack(1)

    @property
    def sum(self) -> np.ndarray:
        return self._stack(2)

    def refresh_max(self) -> np.ndarray:
        return self.max

    @property
    def sliding_window(self) -> np.ndarray:
        """The (n, H, W) ring in the reference's slot order (utils.py:263-265, :276-281); a read-back copy."""
        return _window(self._eng, self.n)

    @property
    def std(self) -> float:
        """utils.py:309-321, force_int branch."""
        assert self.calc_std, "calc_std should be applied when initialized."
        v = C.c_double()
        check(self._eng.lib.mdb_get_std(self._eng.handle, C.byref(v)), "SlidingWindow.std")
        return np.float64(v.value)


def _window(eng: "_Engine", n: int) -> np.ndarray:
    out = np.empty((n, eng.H, eng.W), np.uint8)
    check(eng.lib.mdb_get_window(eng.handle, _ptr(out), 0), "sliding_window")
    return out


class SNR_SW(object):
    """View of the detector's main window with the attributes of SNR_SW (MetLib/Detector.py:34-127)
    that callers read: `snr`, `std_roi`, `n`, `timer`, `length`, `max`, `mean`, `sum`."""

    def __init__(self, det: "LineDetector") -> None:
        self._det = det
        self.n = det.stack_maxsize
        self.std_roi = det._roi
        self.est_snr = True
        self.nz_interval = det.bi_cfg.interval
        self.std_interval = self.nz_interval * self.n

    @property
    def timer(self) -> int:
        return self._det._timer

    @property
    def length(self) -> int:
        return min(self.n, self._det._timer)

    @property
    def snr(self) -> float:
        return self._det._snr

    def _stack(self, which: int):
        eng = self._det._eng
        out = np.empty((eng.H, eng.W), np.uint32 if which == 2 else np.uint8)
        args = [None, None, None]
        args[which] = _ptr(out)
        check(eng.lib.mdb_get_stack(eng.handle, *args), "SNR_SW")
        return out

    @property
    def max(self):
        return self._stack(0)

    @property
    def mean(self):
        return self._stack(1)

    @property
    def sum(self):
        return self._stack(2)

    @property
    def sliding_window(self) -> np.ndarray:
        """utils.py:263-265: the (n, H, W) ring in the reference's slot order (a read-back copy)."""
        return _window(self._det._eng, self.n)


class BaseDetector(metaclass=ABCMeta):
    """BaseDetector (MetLib/Detector.py:130-157)."""

    @abstractmethod
    def __init__(self, *args: Any) -> None:
        pass

    @abstractmethod
    def update(self, new_frame: np.ndarray) -> None:
        pass

    @abstractmethod
    def detect(self):
        pass

    def visu(self) -> list:
        return []


class LineDetector(BaseDetector):
    """LineDetector (MetLib/Detector.py:160-242): window, adaptive threshold policy, dynamic mask.

    Extra keyword arguments (not in the reference): `device` (CUDA ordinal), `max_batch` (capacity of
    `detect_many`), `apply_mask` (multiply frames by `mask` on the device, i.e. do the loader's
    Transform.mask_with, MetLib/imgproc.py:96-101, inside the library)."""
    abs_sensitivity = {"high": 3, "normal": 5, "low": 7}

    def __init__(self, window_sec: float, fps: float, mask: np.ndarray, num_cls: int,
                 cfg: BinaryCfg, logger: Any = None, *, device: int = 0, max_batch: int = 1,
                 apply_mask: bool = False, _detector: int = 0):
        self.mask = mask
        self.num_cls = num_cls
        self.logger = logger
        self.mask_area = np.sum(self.mask)
        self.bi_cfg = cfg.binary
        self.hough_cfg = cfg.hough_line
        self.dynamic_cfg = cfg.dynamic
        self.stack_maxsize = int(window_sec * fps)
        if self.bi_cfg.adaptive_bi_thre:
            self.bi_threshold = self.abs_sensitivity[self.bi_cfg.sensitivity]
        else:
            self.bi_threshold = self.bi_cfg.init_value
        self.bi_threshold_float = self.bi_threshold
        self.max_allow_gap = 0.05
        self._roi = select_subarea(np.asarray(mask), self.bi_cfg.area)
        self._eng = _Engine(mask, self.stack_maxsize, adaptive=self.bi_cfg.adaptive_bi_thre,
                            init_value=self.bi_cfg.init_value, sensitivity=self.bi_cfg.sensitivity,
                            interval=self.bi_cfg.interval, roi=self._roi,
                            hough=(self.hough_cfg.threshold, self.hough_cfg.min_len,
                                   self.hough_cfg.max_gap),
                            dy_mask=self.dynamic_cfg.dy_mask, max_batch=max_batch, device=device,
                            apply_mask=apply_mask, detector=_detector)
        self._timer = 0
        self._snr = 0
        self.stack = SNR_SW(self)

    def detect(self):
        return [], []

    def update(self, new_frame: np.ndarray) -> None:
        f = self._eng.check_frame(new_frame)
        check(self._eng.lib.mdb_update(self._eng.handle, _ptr(f), 0), "update")
        self._timer += 1

    def visu(self):
        return super().visu()

    def close(self):
        self._eng.close()


class ClassicDetector(LineDetector):
    """ClassicDetector (MetLib/Detector.py:245-299): 4-frame window, frame absdiff -> threshold -> dilate
    -> invert, bitwise-and, second absdiff -> threshold -> dilate, HoughLinesP with the configured
    maxLineGap; every raw segment is returned, cls_pred[:, 0] = 1 (no NMS).  `detect_many` is the
    batched form.  Kernels: csrc/classic.cuh."""
    classic_max_size = 4

    def __init__(self, window_sec: float, fps: float, mask: np.ndarray, num_cls: int,
                 cfg: BinaryCfg, logger: Any = None, **kw):
        window_sec = self.classic_max_size / fps  # Detector.py:254: the argument is ignored
        if int(window_sec * fps) != self.classic_max_size:
            raise NotImplementedError(f"fps={fps}: int((4/fps)*fps) != 4 -- the reference's window collapses to 3 "
                                      "frames there and its frame indices alias (Detector.py:254-261); not reproduced")
        super().__init__(window_sec, fps, mask, num_cls, cfg, logger, _detector=1, **kw)
        self.linesp_ext = []
        self._dst_cache = None

    def _result(self, i: int):
        eng = self._eng
        fi = eng.infos[i]
        self.bi_threshold = fi.bi_threshold
        self.bi_threshold_float = fi.bi_threshold_float
        self._snr = fi.snr
        if fi.timer < self.stack_maxsize:
            return [], []  # fewer than four frames: LineDetector.detect() (Detector.py:222-223, :264-265)
        if fi.lines_num > MAX_LINES:
            # the reference returns every segment (Detector.py:282-292): fetch the frame's full list
            lines = np.empty((fi.lines_num, 4), np.int32)
            got = C.c_int32()
            check(eng.lib.mdb_get_raw_lines(eng.handle, i, _ptr(lines), fi.lines_num, C.byref(got)), "raw lines")
        else:
            lines = eng.raw[i, :fi.n_raw].copy() if fi.n_raw else []
        cls_pred = np.zeros((len(lines), self.num_cls))
        cls_pred[:, 0] = 1
        return lines, cls_pred

    def detect(self):
        eng = self._eng
        check(eng.lib.mdb_detect(eng.handle, C.byref(eng.infos[0]), _ptr(eng.lines), _ptr(eng.prob),
                                 _ptr(eng.raw)), "detect")
        self._dst_cache = None
        lines, cls_pred = self._result(0)
        self.linesp_ext = lines
        return lines, cls_pred

    @property
    def dst(self) -> np.ndarray:
        """The thresholded, dilated difference mask HoughLinesP saw (a local of detect() in the
        reference, Detector.py:274-281; zeros before the fourth frame)."""
        if self._dst_cache is None:
            out = np.empty((self._eng.H, self._eng.W), np.uint8)
            check(self._eng.lib.mdb_get_dst(self._eng.handle, _ptr(out), 0), "dst")
            self._dst_cache = out
        return self._dst_cache

    def detect_many(self, frames, *, on_device: bool = False, return_dst: bool = False):
        """`[ (self.update(f), self.detect())[1] for f in frames ]` in one library call."""
        eng = self._eng
        if on_device:
            ptr, T = frames
        else:
            frames = eng.check_frame(np.asarray(frames))
            if frames.ndim != 3:
                raise ValueError("frames must be (T, H, W)")
            T, ptr = len(frames), frames.ctypes.data
        if T == 0:
            return []
        if T > eng.max_batch:
            raise ValueError(f"T={T} exceeds max_batch={eng.max_batch} given at construction")
        dst_out = np.empty((T, eng.H, eng.W), np.uint8) if return_dst else None
        check(eng.lib.mdb_detect_batch(eng.handle, ptr, T, int(on_device), C.byref(eng.infos), _ptr(eng.lines),
                                       _ptr(eng.prob), _ptr(eng.raw), _ptr(dst_out), 0), "detect_many")
        self._timer += T
        self._dst_cache = None
        self.last_infos = [dict(timer=fi.timer, bi_threshold=fi.bi_threshold, n_on=fi.n_on,
                                bi_threshold_float=fi.bi_threshold_float, snr=fi.snr, lines_num=fi.lines_num)
                           for fi in eng.infos[:T]]
        res = [self._result(i) for i in range(T)]
        self.linesp_ext = res[-1][0]
        return (res, dst_out) if return_dst else res

    def visu(self):
        raise NotImplementedError  # as the reference (Detector.py:298-299)


class M3Detector(LineDetector):
    """M3Detector (MetLib/Detector.py:302-448): max-minus-mean over the window, median, threshold,
    close, dynamic mask, probabilistic Hough, NMS.  `update`/`detect` are the reference's per-frame
    calls; `detect_many` is the batched form (T x {update; detect}) that the throughput path uses."""

    def __init__(self, window_sec: float, fps: float, mask: np.ndarray, num_cls: int,
                 cfg: BinaryCfg, logger: Any = None, **kw):
        super().__init__(window_sec, fps, mask, num_cls, cfg, logger, **kw)
        self.lines_num = 0
        self.filtered_line_num = 0
        self.dst_sum = 0.0
        self.linesp_ext = np.array([])
        self._dst_cache = None

    # -- reference per-frame API ---------------------------------------------------------------
    def detect(self):
        eng = self._eng
        check(eng.lib.mdb_detect(eng.handle, C.byref(eng.infos[0]), _ptr(eng.lines), _ptr(eng.prob),
                                 _ptr(eng.raw)), "detect")
        self._dst_cache = None
        return self._unpack(0)

    def _unpack_all(self, T: int, build: bool = True):
        """Results of a finished batch as a list of (lines, cls_pred).  The rows of all frames are gathered in two
        vectorised passes; a frame's result is a pair of views into those arrays (frames without lines share one empty
        pair), so the per-frame Python work is two slices.  build=False: only the bookkeeping (tie re-NMS, last_infos)."""
        eng = self._eng
        info = np.frombuffer(eng.infos, dtype=_INFO_DTYPE, count=T)
        tied = np.nonzero(info["len_ties"])[0]
        if len(tied):
            self._renms_tied(info, tied)
        out = None
        if build:
            nl = info["n_lines"]
            empty = (np.array([]), np.zeros((0, self.num_cls)))
            out = [empty] * T
            idx = np.nonzero(nl)[0]
            if len(idx):
                cnt = nl[idx].astype(np.int64)
                off = np.concatenate(([0], np.cumsum(cnt)))
                rows = np.repeat(idx, cnt)
                cols = np.arange(int(off[-1])) - np.repeat(off[:-1], cnt)
                lines_all = eng.lines[rows, cols]
                p = eng.prob[rows, cols]
                cls_all = np.zeros((len(p), self.num_cls))
                cls_all[:, -1] = p
                cls_all[:, 0] = 1 - p
                offl = off.tolist()
                for j, i in enumerate(idx.tolist()):
                    out[i] = (lines_all[offl[j]:offl[j + 1]], cls_all[offl[j]:offl[j + 1]])
        self._unpack(T - 1)
        self.last_infos = info.copy()
        return out

    def _renms_tied(self, info, tied):
        """Batch form of _renms_if_tied: the squared lengths of all tied frames' segments in one pass (the reference's
        int32 expression, utils.py:802-803), one np.argsort per frame on exactly the array the reference would sort,
        one library call for the greedy passes."""
        eng = self._eng
        nraw = info["n_raw"][tied].astype(np.int64)
        off = np.concatenate(([0], np.cumsum(nraw)))
        rows = np.repeat(tied, nraw)
        cols = np.arange(int(off[-1])) - np.repeat(off[:-1], nraw)
        seg = eng.raw[rows, cols]
        l2 = np.power(seg[:, 3] - seg[:, 1], 2) + np.power(seg[:, 2] - seg[:, 0], 2)
        orders = np.empty(len(l2), np.int32)
        for j in range(len(tied)):
            a, b = off[j], off[j + 1]
            orders[a:b] = np.argsort(l2[a:b])[::-1]
        fr = np.ascontiguousarray(tied, np.int32)
        check(eng.lib.mdb_lineset_nms_frames(len(fr), _ptr(fr), _ptr(orders), _ptr(off), _ptr(eng.raw), C.byref(eng.infos),
                                             _ptr(eng.lines), _ptr(eng.prob)), "lineset_nms")

    def _renms_if_tied(self, i: int, info=None):
        """The library orders equal-length segments by descending index.  The reference's order is
        `np.argsort(length_sqr)[::-1]` (utils.py:804), and among EQUAL lengths numpy's result is its own business: an
        insertion sort (descending index after the reversal) on some builds / CPUs, a SIMD sorting network with another
        tie order on others (x86-simd-sort on AVX-512 / AVX2 hosts), whatever the segment count.  So every frame whose
        raw segments tie in length has its NMS redone with numpy's order on this host (lineset_nms above): the result
        is what the reference returns on the machine it runs on."""
        eng = self._eng
        fi = eng.infos[i]
        if not fi.len_ties:  # set by the library's NMS pass (mdb_frame_info.len_ties)
            return
        raw = eng.raw[i, :fi.n_raw]
        lines, prob = lineset_nms(raw)
        k = len(lines)
        eng.lines[i, :k] = lines
        eng.prob[i, :k] = prob
        fi.n_lines = k
        if info is not None:
            info["n_lines"][i] = k

    def _unpack(self, i: int):
        eng = self._eng
        self._renms_if_tied(i)
        fi = eng.infos[i]
        self.bi_threshold = fi.bi_threshold
        self.bi_threshold_float = fi.bi_threshold_float
        self._snr = fi.snr
        self.dst_sum = fi.dst_sum
        self.gap = fi.gap
        self.lines_num = fi.lines_num
        self.filtered_line_num = fi.n_lines
        self.n_on = fi.n_on
        self.linesp_ext = eng.raw[i, :fi.n_raw].copy() if fi.n_raw else np.array([])
        k = fi.n_lines
        if k > 0:
            lines = eng.lines[i, :k].copy()
            p = eng.prob[i, :k]
            cls_pred = np.zeros((k, self.num_cls))
            cls_pred[:, -1] = p
            cls_pred[:, 0] = 1 - p
        else:
            lines = np.array([])
            cls_pred = np.zeros((0, self.num_cls))
        return lines, cls_pred

    @property
    def dst(self) -> np.ndarray:
        """Binary mask of the most recent detect (Detector.py:371), fetched on demand."""
        if self._dst_cache is None:
            eng = self._eng
            out = np.empty((eng.H, eng.W), np.uint8)
            check(eng.lib.mdb_get_dst(eng.handle, _ptr(out), 0), "dst")
            self._dst_cache = out
        return self._dst_cache

    # -- batched API ---------------------------------------------------------------------------
    def detect_many(self, frames, *, on_device: bool = False, return_dst: bool = False,
                    dst_out: Optional[np.ndarray] = None):
        """Equivalent to `[ (self.update(f), self.detect())[1] for f in frames ]` in ONE library call.

        frames: (T,H,W) uint8 numpy array (host), or -- with on_device=True -- an int device pointer
        wrapped as (ptr, T).  Returns a list of (lines, cls_pred) per frame; per-frame scalars are in
        `self.last_infos` (structured array: timer, bi_threshold, n_on, bi_threshold_float, snr, dst_sum,
        gap, lines_num, n_raw, n_lines), raw Hough segments in `self.last_raw[i]`.  With return_dst the masks
        come back as a (T,H,W) array as second item.  Only frames that have lines cost per-frame Python work."""
        eng = self._eng
        if on_device:
            ptr, T = frames
        else:
            frames = eng.check_frame(np.asarray(frames))
            if frames.ndim != 3:
                raise ValueError("frames must be (T, H, W)")
            T, ptr = len(frames), frames.ctypes.data
        if T == 0:
            return []
        if T > eng.max_batch:
            raise ValueError(f"T={T} exceeds max_batch={eng.max_batch} given at construction")
        dst = None
        if return_dst and dst_out is None:
            dst_out = np.empty((T, eng.H, eng.W), np.uint8)
        if dst_out is not None:
            dst = dst_out.ctypes.data
        check(eng.lib.mdb_detect_batch(eng.handle, ptr, T, int(on_device), C.byref(eng.infos),
                                       _ptr(eng.lines), _ptr(eng.prob), _ptr(eng.raw), dst, 0),
              "detect_many")
        self._timer += T
        self._dst_cache = None
        res = self._unpack_all(T)  # also sets last_infos (structured array, one record per frame)
        nraw = self.last_infos["n_raw"].astype(np.int64)
        rows = np.repeat(np.arange(T), nraw)
        cols = np.arange(int(nraw.sum())) - np.repeat(np.cumsum(nraw) - nraw, nraw)
        self.last_raw = _RaggedRows(eng.raw[rows, cols].reshape(-1, 4), np.concatenate(([0], np.cumsum(nraw))))
        if dst_out is not None:
            return res, dst_out
        return res

    def submit(self, ptr: int, T: int, on_device: bool):
        """Asynchronous half of detect_many (mdb_submit_batch): enqueues the copy (host input) and all
        kernels of one batch and returns.  Up to three batches may be in flight, so the host->device
        copy of the next batch overlaps the kernels of this one.  Host buffers should be pinned and
        must stay untouched until the matching collect(); device buffers likewise."""
        check(self._eng.lib.mdb_submit_batch(self._eng.handle, ptr, T, int(on_device)), "submit")
        self._timer += T
        if not hasattr(self, "_pending"):
            self._pending = []
        self._pending.append(T)

    def submit_thr(self, ptr: int, T: int, on_device: bool, thr: np.ndarray, thr_f: np.ndarray, snr: np.ndarray,
                   halo: bool = False):
        """submit() with the per-frame thresholds supplied by the caller instead of the detector's own noise / EMA
        recurrence (mdb_submit_batch_thr): the form a rank of a time-sharded run uses (sharding.py).  The three
        arrays (int32, float64, float64; T entries) are copied before the call returns.  halo=True marks look-back
        frames whose results are not wanted (MDB_SUBMIT_HALO: window and dynamic-mask history only)."""
        thr = np.ascontiguousarray(thr, np.int32)
        thr_f = np.ascontiguousarray(thr_f, np.float64)
        snr = np.ascontiguousarray(snr, np.float64)
        if not (len(thr) == len(thr_f) == len(snr) == T):
            raise ValueError("threshold arrays must have T entries")
        check(self._eng.lib.mdb_submit_batch_ex(self._eng.handle, ptr, T, int(on_device), _ptr(thr), _ptr(thr_f),
                                                _ptr(snr), 1 if halo else 0), "submit_thr")
        self._timer += T
        if not hasattr(self, "_pending"):
            self._pending = []
        self._pending.append(T)

    def seek(self, timer: int):
        """Start the frame counter at a global frame index (mdb_seek; only before the first frame / after reset())."""
        check(self._eng.lib.mdb_seek(self._eng.handle, int(timer)), "seek")
        self._timer = int(timer)

    def reset(self):
        """Back to the freshly constructed state, keeping the device buffers (mdb_reset)."""
        check(self._eng.lib.mdb_reset(self._eng.handle), "reset")
        self._timer = 0
        self._pending = []
        self._dst_cache = None

    def noise_sums_device(self, segments) -> list:
        """Integer noise sums of the sample timers among device frames (mdb_noise_sums_dev), several segments in one
        call.  segments: iterable of (ptr, T, t0); returns one (T, 2) uint64 array per segment (zero rows for frames
        that are no sample or whose window is not inside the segment)."""
        segs = list(segments)
        if not segs:
            return []
        ptrs = (C.c_void_p * len(segs))(*[int(s[0]) for s in segs])
        Ts = np.array([int(s[1]) for s in segs], np.int32)
        t0s = np.array([int(s[2]) for s in segs], np.int64)
        sums = np.zeros((int(Ts.sum()), 2), np.uint64)
        check(self._eng.lib.mdb_noise_sums_dev(self._eng.handle, len(segs), C.cast(ptrs, C.c_void_p), _ptr(Ts), _ptr(t0s),
                                               _ptr(sums)), "noise_sums_device")
        offs = np.concatenate(([0], np.cumsum(Ts)))
        return [sums[offs[k]:offs[k + 1]] for k in range(len(segs))]

    def collect(self, want_lines: bool = True, want_infos: bool = False):
        """Waits for the oldest submitted batch (mdb_collect_batch); returns its list of
        (lines, cls_pred), or None with want_lines=False (scalars of the last frame still update; with want_infos the
        per-frame records `last_infos` and the NMS rows in the engine's arrays are complete too -- the form a caller
        takes that packs line records itself, sharding.pack_line_records)."""
        eng = self._eng
        T = self._pending.pop(0)
        check(eng.lib.mdb_collect_batch(eng.handle, C.byref(eng.infos), _ptr(eng.lines), _ptr(eng.prob),
                                        _ptr(eng.raw), None, 0), "collect")
        self._dst_cache = None
        if not want_lines and not want_infos:
            self._unpack(T - 1)
            return None
        return self._unpack_all(T, build=want_lines)

    def visu(self):
        """Reference visu() (Detector.py:394-448) needs MetLib.metvisu; without it: no overlays."""
        try:
            from MetLib.metvisu import (DrawRectVisu, ImgVisuAttrs, SquareColorPair,  # type: ignore
                                        TextColorPair, TextVisu)
        except Exception:
            return []
        x1, y1, x2, y2 = self.stack.std_roi
        return [
            ImgVisuAttrs("mix_bg", img=self.dst // 255, weight=0.5, color="yellow"),
            TextVisu("std_value", position="left-top", color="green",
                     text_list=[TextColorPair(text=f"STD:{self.stack.snr:.4f}")]),
            TextVisu("bi_value", position="left-top", color="green", text_list=[TextColorPair(
                text=f"Bi_Threshold: {self.bi_threshold} (rounded from {self.bi_threshold_float:.4f})")]),
            TextVisu("lines_num", position="left-top", color="green", text_list=[TextColorPair(
                text=f"Line num: {self.lines_num} (filtered: {self.filtered_line_num})")]),
            TextVisu("area_ratio", position="left-top", color="green",
                     text_list=[TextColorPair(text=f"Diff Area: {self.dst_sum:.2f}%")]),
            TextVisu("lines_warning", position="left-top", color="red", text_list=[TextColorPair(
                text="WARNING: TOO MANY LINES!" if self.lines_num > 10 else "")]),
            DrawRectVisu("std_roi_area", pair_list=[SquareColorPair(dot_pair=([y1, x1], [y2, x2]))],
                         color="purple"),
        ]
