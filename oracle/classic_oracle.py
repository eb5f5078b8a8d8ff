"""CPU restatement of the reference's ClassicDetector -- TEST INFRASTRUCTURE ONLY (only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import it).

MetLib/Detector.py:245-299 (ClassicDetector: 4-frame window, frame absdiff -> threshold -> dilate ->
invert, bitwise-and, second absdiff -> threshold -> dilate, cv2.HoughLinesP with the configured,
NOT adaptive, maxLineGap; every raw segment is returned with cls_pred[:, 0] = 1, no NMS) on top of
LineDetector.__init__/update (:186-229: SNR_SW noise estimate over the 4-frame window, adaptive
threshold).  Pinned on golden vectors from the live reference (tests/golden/classic_*.npz).
backend "cv2" = the reference's own call sites; "numpy" = the restatements in m3_oracle.
"""
from __future__ import annotations

import numpy as np

from . import m3_oracle as O

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

CLASSIC_MAX_SIZE = 4  # Detector.py:249


class ClassicDetectorOracle:
    def __init__(self, window_sec: float, fps: float, mask: np.ndarray, num_cls: int = 10, *,
                 adaptive=True, init_value=7, sensitivity="normal", area=0.1, interval=2,
                 hough=(10, 10, 10), backend="cv2"):
        if backend == "cv2" and cv2 is None:
            raise RuntimeError("cv2 backend requested but cv2 is not importable")
        window_sec = CLASSIC_MAX_SIZE / fps  # Detector.py:254 (the argument is ignored)
        self.backend = backend
        self.mask = mask
        self.num_cls = num_cls
        self.adaptive, self.sens = adaptive, sensitivity
        self.h_thr, self.h_min_len, self.h_max_gap = hough
        self.stack_maxsize = int(window_sec * fps)
        self.stack = O.SNRSlidingWindow(self.stack_maxsize, mask, area, interval)
        self.bi_threshold = O._ABS_SENS[sensitivity] if adaptive else init_value
        self.bi_threshold_float = self.bi_threshold
        self.kernel = np.ones((3, 3), np.uint8)
        self.linesp_ext = []
        self.dst = None

    def update(self, frame: np.ndarray):  # LineDetector.update, Detector.py:225-229
        self.stack.update(frame)
        if self.adaptive and self.stack.snr != 0:
            self.bi_threshold_float = O._SENS[self.sens](self.stack.snr)
            self.bi_threshold = round(self.bi_threshold_float)

    def _thr_dilate(self, d):
        if self.backend == "cv2":
            _, b = cv2.threshold(d, self.bi_threshold, 255, cv2.THRESH_BINARY)
            return cv2.dilate(b, self.kernel)
        return O.dilate3(O.threshold_binary(d, self.bi_threshold))

    def detect(self):  # Detector.py:257-296
        id3, id2, id1, id0 = [self.stack.cur_index - i for i in range(CLASSIC_MAX_SIZE)]
        sw = self.stack.sliding_window
        if self.stack.timer < self.stack_maxsize:
            self.dst = None
            self.linesp_ext = []
            return [], []
        absdiff = (lambda a, b: cv2.absdiff(a, b)) if self.backend == "cv2" else \
            (lambda a, b: np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.uint8))
        diff23 = 255 - self._thr_dilate(absdiff(sw[id2], sw[id3]))
        f1 = np.bitwise_and(diff23, sw[id0])
        f2 = np.bitwise_and(diff23, sw[id1])
        dst = self._thr_dilate(absdiff(f1, f2))
        if self.backend == "cv2":
            lp = cv2.HoughLinesP(dst, rho=1, theta=O.PI, threshold=self.h_thr, minLineLength=self.h_min_len,
                                 maxLineGap=self.h_max_gap)
            lines = [] if lp is None else lp[:, 0, :]
        else:
            lines = O.hough_lines_p(dst, self.h_thr, self.h_min_len, self.h_max_gap)
        self.linesp_ext = lines
        self.dst = dst
        cls_pred = np.zeros((len(lines), self.num_cls))
        cls_pred[:, 0] = 1
        return lines, cls_pred
