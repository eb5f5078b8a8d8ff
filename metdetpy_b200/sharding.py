"""Time-sharding of one frame stream over several detectors (one per GPU) with exact results.

The reference is one sequential loop (MetDetPy.py:184-227).  Its only cross-frame state is
  * the n-frame window (utils.py:225-321)             -> each shard re-ingests a look-back halo,
  * the dy-mask window over `act` (Detector.py:234-242) -> halo of 2n-2 frames in total,
  * the noise EMA -> threshold (Detector.py:73-91, :225-229), a scalar recurrence over noise samples
    taken at timers 2..n and every interval*n frames.
So: every rank computes the integer noise sums of the sample timers inside its
own chunk (`mdb_noise_sums`), the ranks all-gather those few numbers, every rank replays the scalar
recurrence from t = 0 and gets bit-identical thresholds, then runs its chunk (+ halo, outputs of halo
frames dropped) with those thresholds (`mdb_seek`, `mdb_submit_batch_thr`).  Frames never cross GPUs;
the collectives move O(#samples) and O(#segments) values.  Line records are gathered to rank 0, which
feeds the (sequential, stateful) MeteorCollector in frame order.

The compute engine is injected, so the host logic (planning, exchange, replay, gather) is testable on
CPU with gloo; `CudaEngine` is the product engine.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_SENS = {"low": (2.0, 4.4), "normal": (1.2, 3.6), "high": (0.9, 3.0)}  # Detector.py:177-182
_ABS_SENS = {"high": 3, "normal": 5, "low": 7}  # Detector.py:183


@dataclass
class Shard:
    rank: int
    start: int       # first frame whose result this rank reports
    end: int         # one past the last
    halo_start: int  # first frame this rank ingests (start - (2n-2), clipped at 0)


def plan_shards(total_frames: int, world: int, n: int) -> list[Shard]:
    """Contiguous chunks, as equal as possible; look-back halo of 2n-2 frames."""
    base, rem = divmod(total_frames, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append(Shard(r, s, e, max(0, s - (2 * n - 2))))
        s = e
    return out


def is_noise_sample(tau: int, n: int, interval: int) -> bool:
    """SNR_SW.update's schedule (Detector.py:82-91); tau = SlidingWindow.timer after the update."""
    return (1 < tau <= n) or (tau > n and interval * n > 0 and tau % (interval * n) == 0)


def sigma_from_sums(s1: int, s2: int, L: int, roi_pixels: int) -> float:
    """np.std of the ROI window from exact integer sums -- same operations, in the
    same order, as the device threshold kernel."""
    N = float(L * roi_pixels)
    mean = float(s1) / N
    var = float(s2) / N - mean * mean
    return math.sqrt(var) if var > 0 else 0.0


def replay_thresholds(samples: dict[int, float], total_frames: int, n: int, *, adaptive: bool,
                      init_value: int, sensitivity: str, interval: int):
    """EMA.update (utils.py:334-368) over all noise samples in timer order + LineDetector.update's
    threshold rule (Detector.py:225-229). Returns (thr int32[T], thr_float f64[T], snr f64[T])."""
    m0 = 1.0 - interval / 60.0
    cur_m, warm, t_ema, value = m0, float(n), 0, 0.0
    thr = _ABS_SENS[sensitivity] if adaptive else init_value
    thr_f = float(thr)
    a, b = _SENS[sensitivity]
    out_thr = np.empty(total_frames, np.int32)
    out_f = np.empty(total_frames, np.float64)
    out_snr = np.empty(total_frames, np.float64)
    for i in range(total_frames):
        tau = i + 1
        if is_noise_sample(tau, n, interval):
            if tau not in samples:
                raise KeyError(f"noise sample for timer {tau} is missing")
            sigma = samples[tau]
            if warm != 0.0:
                k = (t_ema * (1.0 - m0)) * warm
                if k < 1.0:
                    u = 1.0 - k
                    cur_m = m0 * (1.0 - u * u)
                else:
                    warm, cur_m = 0.0, m0
            value = cur_m * value + (1.0 - cur_m) * sigma
            t_ema += 1
        if adaptive and value != 0.0:
            thr_f = a * (value * value) + b
            thr = int(round(thr_f))  # Python round(): half to even, as the reference
        out_thr[i], out_f[i], out_snr[i] = thr, thr_f, value
    return out_thr, out_f, out_snr


# ---------------------------------------------------------------------------------------------
class CudaEngine:
    """Product engine: libmetdet_b200 on one GPU."""

    def __init__(self, mask: np.ndarray, n: int, fps: float, cfg, *, device: int = 0, max_batch: int = 64,
                 apply_mask: bool = False):
        from . import _lib
        from .detector import select_subarea
        self._lib = _lib
        self.lib = _lib.load()
        self.mask = np.ascontiguousarray(mask, np.uint8)
        self.n, self.fps, self.cfg, self.device, self.max_batch = n, fps, cfg, device, max_batch
        self.apply_mask = apply_mask
        self.roi = select_subarea(self.mask, cfg.binary.area)
        self.roi_pixels = (self.roi[2] - self.roi[0]) * (self.roi[3] - self.roi[1])

    def noise_sums(self, frames: np.ndarray, t0: int) -> np.ndarray:
        """(T,2) uint64 integer sums for the frames' sample timers (mdb_noise_sums), uploaded in pieces of at most
        max_batch frames plus the n-1 frames of look-back a sample window needs."""
        frames = np.ascontiguousarray(frames, np.uint8)
        T, H, W = frames.shape
        sums = np.zeros((T, 2), np.uint64)
        roi = (C.c_int32 * 4)(*self.roi)
        step = max(self.max_batch, 1)
        for a in range(0, T, step):
            b = min(T, a + step)
            lo = max(0, a - (self.n - 1))
            piece = frames[lo:b]
            part = np.zeros((b - lo, 2), np.uint64)
            self._lib.check(self.lib.mdb_noise_sums(piece.ctypes.data, b - lo, 0, t0 + lo, W, H, self.n,
                                                    int(self.cfg.binary.interval), roi,
                                                    self.mask.ctypes.data if self.apply_mask else None,
                                                    part.ctypes.data, self.device), "mdb_noise_sums")
            sums[a:b] = part[a - lo:]
        return sums

    def _detector(self):
        from .detector import M3Detector
        if getattr(self, "_det", None) is None:
            self._det = M3Detector(self.n / self.fps + 1e-9, self.fps, self.mask, 10, self.cfg, None, device=self.device,
                                   max_batch=self.max_batch, apply_mask=self.apply_mask)
        return self._det

    def close(self):
        if getattr(self, "_det", None) is not None:
            self._det.close()
            self._det = None

    def detect_chunk(self, frames: np.ndarray, t0: int, thr, thr_f, snr, want_dst: bool = False, num_cls: int = 10):
        """Run host frames (global index of frames[0] = t0) with the given per-frame thresholds on one detector that
        is reset and re-seeked per call, batch by batch.  Returns per-frame (lines, cls_pred) and, optionally, the masks."""
        det = self._detector()
        det.num_cls = num_cls
        eng = det._eng
        det.reset()
        det.seek(t0)
        frames = np.ascontiguousarray(frames, np.uint8)
        res, dsts = [], []
        for s in range(0, len(frames), self.max_batch):
            T = min(self.max_batch, len(frames) - s)
            det.submit_thr(frames[s:s + T].ctypes.data, T, False, thr[s:s + T], thr_f[s:s + T], snr[s:s + T])
            det._pending.pop()
            dst = np.empty((T,) + frames.shape[1:], np.uint8) if want_dst else None
            self._lib.check(self.lib.mdb_collect_batch(eng.handle, C.byref(eng.infos), eng.lines.ctypes.data,
                                                       eng.prob.ctypes.data, eng.raw.ctypes.data,
                                                       dst.ctypes.data if want_dst else None, 0), "collect")
            res += [det._unpack(i) for i in range(T)]
            if want_dst:
                dsts.append(dst)
        return res, (np.concatenate(dsts) if want_dst else None)


# ---------------------------------------------------------------------------------------------
def local_samples(engine, frames: np.ndarray, shard: Shard, n: int, interval: int):
    """(timer, sum d, sum d^2) of the noise samples this rank owns: timers in (start, end]."""
    sums = engine.noise_sums(frames, shard.halo_start)
    out = []
    for tau in range(shard.start + 1, shard.end + 1):
        if is_noise_sample(tau, n, interval):
            i = tau - 1 - shard.halo_start
            out.append((tau, int(sums[i, 0]), int(sums[i, 1])))
    return out


def exchange_samples(mine: Sequence[tuple], group=None, device="cpu") -> list[tuple]:
    """all-gather of the ranks' (timer, s1, s2) triples (int64; a few dozen values)."""
    import torch
    import torch.distributed as dist
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return sorted(mine)
    world = dist.get_world_size(group)
    cnt = torch.tensor([len(mine)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cap = max(1, int(max(int(c.item()) for c in cnts)))
    buf = torch.zeros((cap, 3), dtype=torch.int64, device=device)
    if mine:
        buf[:len(mine)] = torch.tensor(mine, dtype=torch.int64, device=device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    out = []
    for c, b in zip(cnts, bufs):
        out += [tuple(int(v) for v in row) for row in b[:int(c.item())].cpu().tolist()]
    return sorted(out)


def gather_lines(per_frame, first_frame: int, group=None, device="cpu"):
    """Gather (frame, x1, y1, x2, y2, nonline_prob) records to rank 0 (others get None).
    per_frame: list of (lines, cls_pred) for frames first_frame, first_frame+1, ..."""
    import torch
    import torch.distributed as dist
    rec = []
    for i, (lines, cls) in enumerate(per_frame):
        rows = np.asarray(lines).reshape(-1, 4)
        if len(rows) == 0:
            continue
        probs = np.asarray(cls).reshape(len(rows), -1)[:, -1]
        for row, p in zip(rows, probs):
            rec.append([first_frame + i, *[int(v) for v in row], float(p)])
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return rec
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cnt = torch.tensor([len(rec)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cap = max(1, int(max(int(c.item()) for c in cnts)))
    buf = torch.zeros((cap, 6), dtype=torch.float64, device=device)
    if rec:
        buf[:len(rec)] = torch.tensor(rec, dtype=torch.float64, device=device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)  # NCCL has no gather-to-one for ragged data; counts trim it
    if rank != 0:
        return None
    out = []
    for c, b in zip(cnts, bufs):
        out += b[:int(c.item())].cpu().tolist()
    out.sort(key=lambda r: r[0])
    return [[int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), r[5]] for r in out]


def detect_sharded(engine, frames: np.ndarray, shard: Shard, total_frames: int, n: int, cfg, *, group=None,
                   device="cpu", want_dst: bool = False):
    """One rank's part of a time-sharded run. `frames` = frames halo_start .. end-1 of the stream.
    Returns (per-frame results for start..end-1, dst or None, gathered records on rank 0)."""
    b = cfg.binary
    mine = local_samples(engine, frames, shard, n, int(b.interval))
    allsum = exchange_samples(mine, group, device)
    samples = {}
    for tau, s1, s2 in allsum:
        samples[tau] = sigma_from_sums(s1, s2, min(n, tau), engine.roi_pixels)
    # a rank only needs thresholds up to its own end; samples of later ranks are ignored by the replay
    thr, thr_f, snr = replay_thresholds({k: v for k, v in samples.items() if k <= shard.end}, shard.end, n,
                                        adaptive=bool(b.adaptive_bi_thre), init_value=int(b.init_value),
                                        sensitivity=b.sensitivity, interval=int(b.interval))
    h0 = shard.halo_start
    res, dst = engine.detect_chunk(frames, h0, thr[h0:shard.end], thr_f[h0:shard.end], snr[h0:shard.end], want_dst)
    skip = shard.start - h0
    res = res[skip:]
    if dst is not None:
        dst = dst[skip:]
    records = gather_lines(res, shard.start, group, device)
    return res, dst, records, (thr[shard.start:shard.end], thr_f[shard.start:shard.end], snr[shard.start:shard.end])


# ---------------------------------------------------------------------------------------------
# Device-resident chunks on the product pipeline (what bench.py --gpus N runs)
# ---------------------------------------------------------------------------------------------
@dataclass
class Segment:
    """T frames, contiguous in device memory at `ptr`, global index of the first one `t0`; `history` frames of the
    stream lie in memory right in front of ptr (needed by noise windows that start before the segment)."""
    ptr: int
    T: int
    t0: int
    history: int = 0


def sample_timers(lo: int, hi: int, n: int, interval: int) -> np.ndarray:
    """The noise-sample timers tau with lo < tau <= hi (SNR_SW.update's schedule, Detector.py:82-91), ascending."""
    parts = [np.arange(max(lo + 1, 2), min(hi, n) + 1, dtype=np.int64)]
    si = interval * n
    if si > 0:
        first = max(lo + 1, n + 1)
        parts.append(np.arange(-(-first // si) * si, hi + 1, si, dtype=np.int64))
    return np.concatenate(parts)


def chunk_noise_samples(det, segments: Sequence[Segment], n: int, interval: int, start: int, end: int, frame_bytes: int):
    """(timer, sum d, sum d^2) rows (int64 array, ascending timers) of the noise samples with timers in (start, end],
    from device frames: one mdb_noise_sums_dev call for all segments, on the detector's own stream and buffers."""
    taus_all = sample_timers(start, end, n, interval)
    calls, wanted = [], []
    for sg in segments:
        lo, hi = max(sg.t0, start), min(sg.t0 + sg.T, end)
        if hi <= lo:
            continue
        taus = taus_all[np.searchsorted(taus_all, lo, side="right"):np.searchsorted(taus_all, hi, side="right")]
        if len(taus) == 0:
            continue
        back = min(sg.history, n - 1) if sg.t0 > 0 else 0
        first = sg.t0 - back
        if first > 0 and int(taus[0]) - n < first:
            raise ValueError(f"noise sample {int(taus[0])}: its window starts before the frames of the segment")
        calls.append((sg.ptr - back * frame_bytes, sg.T + back, first))
        wanted.append((first, taus))
    out = [np.zeros((0, 3), np.int64)]
    for (first, taus), sums in zip(wanted, det.noise_sums_device(calls)):
        rows = np.empty((len(taus), 3), np.int64)
        rows[:, 0] = taus
        rows[:, 1:] = sums[taus - 1 - first].astype(np.int64)
        out.append(rows)
    return np.concatenate(out)


def replay_thresholds_native(samples: Sequence[tuple], roi_pixels: int, n: int, t_begin: int, t_end: int, *, adaptive: bool,
                             init_value: int, sensitivity: str, interval: int):
    """replay_thresholds in the library's host code (mdb_replay_thresholds): (timer, s1, s2) rows (any order) ->
    (thr int32, thr_float f64, snr f64) for frames t_begin .. t_end-1.  Same arithmetic as the device recurrence."""
    from . import _lib
    lib = _lib.load()
    smp = np.asarray(samples, np.int64).reshape(-1, 3)
    if len(smp) > 1 and np.any(np.diff(smp[:, 0]) < 0):
        smp = smp[np.argsort(smp[:, 0], kind="stable")]
    timers = np.ascontiguousarray(smp[:, 0])
    sums = np.ascontiguousarray(smp[:, 1:]).astype(np.uint64)
    m = t_end - t_begin
    thr, thr_f, snr = np.empty(m, np.int32), np.empty(m, np.float64), np.empty(m, np.float64)
    _lib.check(lib.mdb_replay_thresholds(len(smp), timers.ctypes.data, sums.ctypes.data, int(roi_pixels), int(n),
                                         int(interval), int(bool(adaptive)), int(init_value),
                                         {"low": 0, "normal": 1, "high": 2}[sensitivity], int(t_begin), int(t_end),
                                         thr.ctypes.data, thr_f.ctypes.data, snr.ctypes.data), "mdb_replay_thresholds")
    return thr, thr_f, snr


def ragged_index(counts: np.ndarray):
    """(row, column) index arrays that enumerate the first counts[i] entries of every row i, row-major: gathers the
    used part of a padded [T][cap] result array without touching the padding."""
    counts = np.asarray(counts, np.int64)
    total = int(counts.sum())
    rows = np.repeat(np.arange(len(counts)), counts)
    starts = np.cumsum(counts) - counts
    cols = np.arange(total) - np.repeat(starts, counts)
    return rows, cols


def pack_line_records(det, T: int, first_frame: int, out: np.ndarray) -> int:
    """Line records (frame, x1, y1, x2, y2, nonline_prob) of the batch collected last into out[:k] (float64 rows),
    without a per-frame Python loop.  Returns k (clipped to len(out))."""
    eng = det._eng
    nl = det.last_infos["n_lines"][:T]
    rows, cols = ragged_index(nl)
    k = min(len(rows), len(out))
    if k:
        out[:k, 0] = first_frame + rows[:k]
        out[:k, 1:5] = eng.lines[rows[:k], cols[:k]]
        out[:k, 5] = eng.prob[rows[:k], cols[:k]]
    return k


def batch_digest(det, T: int) -> bytes:
    """8-byte digest of the batch collected last, computed by the library while it fills the results (FNV-1a over
    per-frame threshold, on-pixel count, raw segment count and the raw Hough segments themselves).  Two runs that agree
    on every digest produced the same masks' statistics and lines."""
    eng = det._eng
    return int(eng.info("digest_hi")).to_bytes(4, "big") + int(eng.info("digest_lo")).to_bytes(4, "big")


def run_chunk(det, segments: Sequence[Segment], shard: Shard, thr, thr_f, snr, *, thr_base: int = 0, on_batch=None,
              in_flight: int = 3):
    """One rank's chunk on the product pipeline: mdb_reset + mdb_seek(halo_start), then every segment with the
    replayed thresholds (mdb_submit_batch_thr), `in_flight` batches submitted ahead.  `segments` cover
    [halo_start, end) in order; segments that end at or before shard.start are halo (results dropped).
    thr[i] belongs to global frame thr_base + i.  on_batch(det, seg) is called after each collected non-halo batch
    (det.last_infos etc. describe it)."""
    det.reset()
    det.seek(shard.halo_start)
    segs = list(segments)
    if segs and segs[0].t0 != shard.halo_start:
        raise ValueError("segments must start at the shard's halo_start")
    nxt = 0

    def submit(sg):
        a, b = sg.t0 - thr_base, sg.t0 - thr_base + sg.T
        det.submit_thr(sg.ptr, sg.T, True, thr[a:b], thr_f[a:b], snr[a:b], halo=sg.t0 + sg.T <= shard.start)

    for k in range(min(in_flight - 1, len(segs))):
        submit(segs[nxt])
        nxt += 1
    for i, sg in enumerate(segs):
        if nxt < len(segs):
            submit(segs[nxt])
            nxt += 1
        halo = sg.t0 + sg.T <= shard.start
        det.collect(want_lines=False, want_infos=not halo)  # records are packed from the engine's arrays (pack_line_records)
        if not halo and on_batch is not None:
            on_batch(det, sg)
