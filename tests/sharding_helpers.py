"""Test-only compute engine for metdetpy_b200.sharding built on the CPU oracle (tests may use it)."""
import numpy as np

from metdetpy_b200 import sharding as S
from oracle import m3_oracle as O


class OracleEngine:
    def __init__(self, mask, n, fps, cfg):
        self.mask, self.n, self.fps, self.cfg = mask, n, fps, cfg
        self.roi = O.select_subarea(mask, cfg.binary.area)
        self.roi_pixels = (self.roi[2] - self.roi[0]) * (self.roi[3] - self.roi[1])

    def noise_sums(self, frames, t0):
        r0, c0, r1, c1 = self.roi
        T = len(frames)
        out = np.zeros((T, 2), np.uint64)
        for i in range(T):
            tau = t0 + i + 1
            if not S.is_noise_sample(tau, self.n, int(self.cfg.binary.interval)):
                continue
            L = min(self.n, tau)
            if t0 != 0 and tau < t0 + self.n:
                continue
            win = frames[i - L + 1:i + 1, r0:r1, c0:c1].astype(np.int64)
            sx, sxx = win.sum(0), (win * win).sum(0)
            m = sx // L
            out[i, 0] = int((sx - L * m).sum())
            out[i, 1] = int((sxx - 2 * m * sx + L * m * m).sum())
        return out

    def detect_chunk(self, frames, t0, thr, thr_f, snr, want_dst=False):
        b, h, d = self.cfg.binary, self.cfg.hough_line, self.cfg.dynamic
        det = O.M3DetectorOracle(self.n / self.fps + 1e-9, self.fps, self.mask, 10, adaptive=False,
                                 init_value=0, sensitivity=b.sensitivity, area=b.area, interval=b.interval,
                                 hough=(h.threshold, h.min_len, h.max_gap), dy_mask=d.dy_mask, backend="numpy")
        det.stack.timer = det.stack.sub_sw.timer = t0  # global frame counter (mdb_seek)
        if d.dy_mask:
            det.dy_sw.timer = t0
        res, dsts = [], []
        for i, f in enumerate(frames):
            O.SlidingWindow.update(det.stack, f)  # window only: thresholds come from the schedule
            det.bi_threshold = int(thr[i])
            res.append(det.detect())
            dsts.append(det.dst.copy())
        return res, (np.stack(dsts) if want_dst else None)


def sequential_reference(frames, mask, n, fps, cfg):
    b, h, d = cfg.binary, cfg.hough_line, cfg.dynamic
    det = O.M3DetectorOracle(n / fps + 1e-9, fps, mask, 10, adaptive=b.adaptive_bi_thre, init_value=b.init_value,
                             sensitivity=b.sensitivity, area=b.area, interval=b.interval,
                             hough=(h.threshold, h.min_len, h.max_gap), dy_mask=d.dy_mask, backend="numpy")
    thr, snr, dst, lines = [], [], [], []
    for f in frames:
        det.update(f)
        l, c = det.detect()
        thr.append(det.bi_threshold); snr.append(float(det.stack.snr)); dst.append(det.dst.copy())
        lines.append((np.asarray(l).reshape(-1, 4), c))
    return np.array(thr), np.array(snr), np.stack(dst), lines
