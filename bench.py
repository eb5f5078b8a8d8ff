#!/usr/bin/env python
"""bench.py -- frames/s through M3Detector (update+detect) on synthetic 4K streams, B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path (T x {update; detect}: noise sample + threshold recurrence, fused
stack->diff->median->threshold->close->dy-mask chain, PPHT Hough, NMS) over `--bps` batches of `--batch`
synthetic frames per GPU.  Workload at every N: BASELINE.json configs[2] -- synthetic 3840x2160 @30 fps,
window n=30, adaptive threshold, dynamic mask, HoughLinesP.

N = 1: one sequential stream on the product pipeline (three batches in flight).
N > 1: ONE stream of N time chunks (weak scaling: K steps per rank) through the time-sharded algorithm of
SURVEY 8(e) on the same pipeline: every rank computes the integer noise sums of its chunk (mdb_noise_sums_dev), one
NCCL all-gather of the (timer, sum d, sum d^2) triples, bit-identical threshold replay (mdb_replay_thresholds),
mdb_seek + a (2n-2)-frame halo batch + the chunk with replayed thresholds (mdb_submit_batch_thr); line records are
all-gathered once per step on a side stream.  No frame crosses GPUs.  Every rank then re-runs frames 0 .. end of its
chunk sequentially and compares per-batch digests (thresholds, on-pixel counts, raw Hough segments): "sharded_parity".

Prints ONE JSON line (see README / DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "4K frames/sec through M3Detector.detect()"
UNIT = "frames/s"
DISTINCT = 4  # resident batches of the periodic synthetic stream


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--fps", type=float, default=30.0)
    ap.add_argument("--window", type=int, default=30)
    ap.add_argument("--batch", type=int, default=512, help="frames per library call (per GPU)")
    ap.add_argument("--bps", type=int, default=16, help="batches per step (a multiple of 4 keeps every rank's chunk in phase "
                                                        "with the replayed stream)")
    ap.add_argument("--no-dy", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=40, help="frames in the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the side measurements of the SURVEY 8f rows")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (2, 4, 5)")
    ap.add_argument("--no-extra", action="store_true", help="skip per_frame_api and dense_regime")
    ap.add_argument("--no-parity", action="store_true", help="skip the sequential re-run that checks the sharded result")
    ap.add_argument("--parity-budget", type=int, default=6000, help="most batches a rank re-runs sequentially for the parity check")
    ap.add_argument("--generic-kernel", action="store_true", help="force the per-frame fused kernel")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic {a.width}x{a.height} @{a.fps:g}fps, window={a.window}, adaptive threshold, "
            f"dy_mask={'off' if a.no_dy else 'on'}, HoughLinesP(10,10,10)")


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms` process streams samples; nothing runs in this process while timing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        if os.environ.get("BENCH_NO_SMI"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", os.environ.get("BENCH_SMI_MS", "100")], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            self.rows.append([c.strip() for c in line.split(",")])

    def summary(self):
        def num(v):
            try:
                return float(v)
            except Exception:
                return None
        rows = [r for r in self.rows if len(r) >= 9]
        pw = [num(r[3]) for r in rows if num(r[3]) is not None]
        # "under load": samples whose power draw is above the midpoint of the observed range
        thr = (min(pw) + max(pw)) / 2 if pw else 0
        load = [r for r in rows if num(r[3]) is not None and num(r[3]) >= thr] or rows
        sm = [num(r[1]) for r in load if num(r[1]) is not None]
        mx = [num(r[2]) for r in rows if num(r[2]) is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows), "power_w_max": max(pw) if pw else None}


def make_cfg(dy=True, adaptive=True, init_value=7):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    return BinaryCfg(BinaryCoreCfg(adaptive, init_value, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(dy, 5))


def cpu_reference_detector(W, H, n, fps, dy=True, adaptive=True, init_value=7, mask=None):
    """The reference's CPU implementation of the path: the oracle in its cv2 backend executes the same
    numpy + cv2 calls, in the same order, as MetLib/Detector.py (the reference is pure Python and
    /root/reference does not exist on the GPU box)."""
    from oracle import m3_oracle as O
    import cv2
    mask = np.ones((H, W), np.uint8) if mask is None else mask
    det = O.M3DetectorOracle(n / fps + 1e-9, fps, mask, 10, adaptive=adaptive, init_value=init_value,
                             sensitivity="normal", area=0.1, interval=2, hough=(10, 10, 10),
                             dy_mask=dy, backend="cv2")
    return det, cv2.getNumThreads()


def run_cpu_sample(a, nframes):
    """Bounded CPU sample: fill the window (untimed), then time `nframes` x (update; detect)."""
    from metdetpy_b200 import synth
    det, threads = cpu_reference_detector(a.width, a.height, a.window, a.fps, dy=not a.no_dy)
    sky = synth.make_sky(a.width, a.height)
    warm = a.window + 2
    frames = [synth.make_frame(t, sky, a.width, a.height, a.fps) for t in range(warm + nframes)]
    for t in range(warm):
        det.update(frames[t]); det.detect()
    t0 = time.perf_counter()
    for t in range(warm, warm + nframes):
        det.update(frames[t]); det.detect()
    dt = time.perf_counter() - t0
    return nframes / dt, threads, warm


# ---------------------------------------------------------------------------------------------
class Stream:
    """Periodic synthetic stream resident on the device: frame t has the content of frame t mod P (P = DISTINCT * B),
    generated with that period (synth.make_stream_device).  In front of the period sit `pad` frames that repeat its
    end, so that every batch and every look-back halo / noise window is one contiguous block of memory."""

    def __init__(self, B, W, H, fps, dev, pad, distinct=DISTINCT, quiet=0, t_base=0):
        import torch
        from metdetpy_b200 import synth
        self.B, self.W, self.H, self.pad, self.P = B, W, H, pad, distinct * B
        self.HW = W * H
        self.buf = torch.empty((pad + self.P, H, W), dtype=torch.uint8, device=dev)
        for s in range(distinct):
            self.buf[pad + s * B: pad + (s + 1) * B] = synth.make_stream_device(B, W, H, fps, dev, t0=t_base + s * B,
                                                                                 loop=self.P, quiet=quiet)
        if pad:
            self.buf[:pad] = self.buf[self.P:self.P + pad]
        torch.cuda.synchronize()
        self.base = self.buf.data_ptr()

    def ptr(self, t):
        """device address of global frame t (valid for reads of up to B frames when t % B == 0, and `pad` frames back)"""
        return self.base + (self.pad + t % self.P) * self.HW

    def view(self, t, T):
        i = self.pad + t % self.P
        return self.buf[i:i + T]


def measure_next_rows(a, batch0_ptr, dev, peak):
    """Side measurements of the SURVEY 8(f) rows built so far (not part of the headline value):
    loader preprocessing (4K BGR -> 960x540 gray) and ClassicDetector on the bench's 4K stream."""
    import cv2
    import torch
    from metdetpy_b200.detector import ClassicDetector
    from metdetpy_b200.imgproc import Transform
    from oracle import classic_oracle as CO
    from oracle import preproc_oracle as PO
    out = {}
    # ---- row 1: Transform chain on the device ---------------------------------------------------
    W0, H0, W, H, T = 3840, 2160, 960, 540, 64
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    src = torch.randint(0, 256, (T, H0, W0, 3), dtype=torch.uint8, device=dev, generator=g)
    tr = Transform(device=dev.index or 0)
    tr.opencv_resize([W, H]); tr.opencv_BGR2GRAY(); tr.mask_with(np.ones((H, W), np.uint8))
    ms = []
    for _ in range(8):
        tr.exec_transform_many((src.data_ptr(), T), 1, on_device=True, keep_on_device=True, shape=(H0, W0, 3))
        ms.append(tr.last_kernel_ms())
    ms = float(np.median(ms[2:]))
    rows_used = len(set(np.concatenate(PO.axis_taps(H, H0, False)[:2]).tolist()))
    need = T * rows_used * W0 * 3 + T * H * W
    host = src[:8].cpu().numpy()
    t0 = time.perf_counter()
    for f in host:
        _ = cv2.cvtColor(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    cpu = len(host) / (time.perf_counter() - t0)
    out["preprocess"] = {"workload": f"{T} device-resident {W0}x{H0} BGR frames -> {W}x{H} gray + mask (Transform chain, imgproc.py:82-101)",
                         "frames_per_s": T / (ms * 1e-3), "kernel_ms": ms,
                         "roofline": {"bound": "hbm", "achieved": need / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": need / (ms * 1e-3) / 1e9 / peak,
                                      "bytes": "source rows the taps touch + output frame"},
                         "cpu_baseline": {"value": cpu, "unit": "frames/s", "kind": "reference", "cores": cv2.getNumThreads(),
                                          "sample": "8 frames, cv2.resize + cv2.cvtColor (the reference's own calls)"}}
    tr.close()
    del src
    # ---- row 2: ClassicDetector --------------------------------------------------------------------
    B = min(a.batch, 256)
    mask = np.ones((a.height, a.width), np.uint8)
    # fixed threshold 20: with the adaptive one the 4-frame noise estimate puts the threshold at ~7 grey levels on
    # this sigma=2 stream and frame differences light up ~10 % of the pixels -- a regime in which the reference's
    # own cv2.HoughLinesP needs minutes per 4K frame; the fixed value leaves the streaks (amplitude 60)
    ccfg = make_cfg(dy=False, adaptive=False, init_value=20)
    det = ClassicDetector(1.0, a.fps, mask, 10, ccfg, None, device=dev.index or 0, max_batch=B)
    for _ in range(2):
        det.detect_many((batch0_ptr, B), on_device=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        det.detect_many((batch0_ptr, B), on_device=True)
    dt = time.perf_counter() - t0
    chain_ms, _ = det._eng.fused_time()
    ref = CO.ClassicDetectorOracle(1.0, a.fps, mask, 10, adaptive=False, init_value=20, sensitivity="normal", area=0.1,
                                   interval=2, hough=(10, 10, 10), backend="cv2")
    host = _dev_view(batch0_ptr, 12, a.height, a.width, dev).cpu().numpy()
    for f in host[:4]:
        ref.update(f); ref.detect()
    t0 = time.perf_counter()
    for f in host[4:]:
        ref.update(f); ref.detect()
    cpu = (len(host) - 4) / (time.perf_counter() - t0)
    HW = a.width * a.height
    out["classic_detector"] = {"workload": f"ClassicDetector (Detector.py:245-299), fixed threshold 20, {B} device-resident {a.width}x{a.height} frames per call",
                               "frames_per_s": reps * B / dt, "mask_chain_ms_per_call": chain_ms,
                               "roofline": {"bound": "hbm", "achieved": 2.0 * HW * B / (chain_ms * 1e-3) / 1e9, "peak": peak,
                                            "unit": "GB/s", "frac": 2.0 * HW * B / (chain_ms * 1e-3) / 1e9 / peak,
                                            "bytes": "frame read once + u8 mask written once"},
                               "cpu_baseline": {"value": cpu, "unit": "frames/s", "kind": "port", "cores": cv2.getNumThreads(),
                                                "sample": "8 frames, oracle cv2 backend = the reference's numpy+cv2 calls"}}
    det.close()
    # ---- row 3: MFNR mix stacker (stacker.py:296-403), sigma clipping, device-resident colour clip ------------------
    try:
        import ctypes as C
        from metdetpy_b200 import _lib, stacker
        from metdetpy_b200._lib import check
        from oracle import mfnr_oracle as MO
        Tm, Hm, Wm = 48, 2160, 3840
        g = torch.Generator(device=dev); g.manual_seed(3)
        clip = (torch.randn((Tm, Hm, Wm, 3), device=dev, generator=g) * 5.0 + 50.0).clamp_(0, 255).to(torch.uint8)
        clip[:, 1000:1003, 500:2500] = 240
        torch.cuda.synchronize()
        lib = _lib.load()
        times = []
        for rep in range(3):
            h = C.c_void_p()
            check(lib.mdb_mfnr_create(Hm, Wm, 3, 1, dev.index or 0, C.byref(h)), "mfnr")
            prm = _lib.MfnrParams()
            prm.highlight_preserve, prm.blur_ksize, prm.blur_sigma, prm.bg_algorithm = 0.9, 31, 3.0, 1
            prm.sigma_high = prm.sigma_low = 3.0
            prm.bg_fix_factor, prm.gumbel_mean = 1.5, float(stacker.get_gumbel_mean(Tm))
            outb = torch.empty((Hm, Wm, 3), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            check(lib.mdb_mfnr_reserve(h, Tm), "mfnr")
            for s0 in range(0, Tm, 16):
                check(lib.mdb_mfnr_append(h, clip[s0:s0 + 16].data_ptr(), 16, 1), "mfnr")
            check(lib.mdb_mfnr_finish(h, C.byref(prm), outb.data_ptr(), 1, None), "mfnr")
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            lib.mdb_mfnr_destroy(h)
        dt = float(np.median(times))
        small = clip[:, 540:1080, 960:1920].cpu().numpy()  # a quarter-HD crop of the same clip for the CPU arm
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            MO.mfnr_mix(small, bg_algorithm="sigma-clipping", backend="cv2")
        cpu_dt = (time.perf_counter() - t0) * (Hm * Wm) / (540 * 960)
        out["mfnr_mix_stacker"] = {"workload": f"{Tm} device-resident {Wm}x{Hm} BGR frames, bg sigma-clipping, blur 31 (stacker.py:296-403, connect_lines off)",
                                   "frames_per_s": Tm / dt, "ms_per_clip": dt * 1e3,
                                   "roofline": {"bound": "hbm", "achieved": 3.0 * Tm * Hm * Wm * 3 / dt / 1e9, "peak": peak, "unit": "GB/s",
                                                "frac": 3.0 * Tm * Hm * Wm * 3 / dt / 1e9 / peak,
                                                "bytes": "every frame byte read by the accumulation and by the clipping pass, written once into the resident copy"},
                                   "cpu_baseline": {"value": Tm / cpu_dt, "unit": "frames/s", "kind": "port", "cores": 1,
                                                    "sample": "960x540 crop of the same clip through the oracle (numpy + cv2.GaussianBlur), time scaled by the pixel ratio"}}
        del clip, outb
    except Exception as e:
        out["mfnr_mix_stacker"] = {"error": repr(e)}
    # ---- loader leg: host BGR frames -> Transform chain -> detector, end to end (the reference's main loop from decoded frames)
    try:
        out["loader_leg"] = measure_loader_leg(dev)
    except Exception as e:
        out["loader_leg"] = {"error": repr(e)}
    return out


def measure_loader_leg(dev):
    """What MetDetPy.detect_video does per frame once the decoder has produced it (videoloader.py:300-308, 388;
    MetDetPy.py:192-198): resize -> gray -> mask -> detector.  Source: 4K BGR frames in pinned host memory (NVDEC is not
    reachable from this container: profiles/r02_nvdec_probe2.txt), detector at the reference's 960x540 working size.
    Two preprocessing handles alternate so that the detector of chunk k runs beside the H2D copy of chunk k + 1."""
    import cv2
    import torch
    from metdetpy_b200.detector import M3Detector
    from metdetpy_b200.imgproc import Transform
    W0, H0, W, H, T, n, fps = 3840, 2160, 960, 540, 32, 30, 30.0
    g = torch.Generator(device=dev); g.manual_seed(5)
    hosts = []
    for k in range(2):
        d = (torch.randn((T, H0, W0, 3), device=dev, generator=g) * 4.0 + 45.0).clamp_(0, 255).to(torch.uint8)
        d[:, 700 + 40 * k:704 + 40 * k, 300:1500] = 230
        hb = torch.empty((T, H0, W0, 3), dtype=torch.uint8).pin_memory()
        hb.copy_(d)
        hosts.append(hb)
        del d
    mask = np.ones((H, W), np.uint8)
    trs = []
    for k in range(2):
        tr = Transform(device=dev.index or 0)
        tr.opencv_resize([W, H]); tr.opencv_BGR2GRAY(); tr.mask_with(mask)
        trs.append(tr)
    det = M3Detector(n / fps + 1e-9, fps, mask, 10, make_cfg(), None, device=dev.index or 0, max_batch=T)
    arrs = [h.numpy() for h in hosts]

    def run(nb):
        pending = 0
        for k in range(nb):
            df = trs[k % 2].exec_transform_many(arrs[k % 2], 1, keep_on_device=True)
            det.submit(df.ptr, T, True)
            pending += 1
            if pending == 2:
                det.collect(); pending -= 1
        while pending:
            det.collect(); pending -= 1
    run(3)
    torch.cuda.synchronize()
    nb = 10
    t0 = time.perf_counter()
    run(nb)
    dt = time.perf_counter() - t0
    for tr in trs:
        tr.close()
    det.close()
    # CPU arm: the reference's own calls on the same frames
    cdet, threads = cpu_reference_detector(W, H, n, fps)
    src = arrs[0][:12]
    for f in src[:4]:
        cdet.update(cv2.cvtColor(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) * mask); cdet.detect()
    t0 = time.perf_counter()
    for f in src[4:]:
        cdet.update(cv2.cvtColor(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY) * mask); cdet.detect()
    cpu = 8 / (time.perf_counter() - t0)
    return {"workload": f"{W0}x{H0} BGR frames in pinned host memory -> resize {W}x{H} -> gray -> mask -> M3Detector(window {n}), {T} frames per call",
            "frames_per_s": nb * T / dt, "h2d_gbs": nb * T * H0 * W0 * 3 / dt / 1e9, "h2d_bytes_per_frame": H0 * W0 * 3,
            "bound": "PCIe: every source frame is 24.9 MB",
            "cpu_baseline": {"value": cpu, "unit": "frames/s", "kind": "port", "cores": threads,
                             "sample": "8 frames: cv2.resize + cv2.cvtColor + the oracle's cv2 backend at 960x540 (window not full: its cost does not depend on that)"}}


_VIEWS = {}


def _dev_view(ptr, T, H, W, dev):
    """torch view of T frames at a raw device address inside one of the resident streams"""
    for st in _VIEWS.values():
        off = ptr - st.base
        if 0 <= off < st.buf.numel():
            i = off // st.HW
            return st.buf[i:i + T]
    raise ValueError("address not inside a resident stream")


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from metdetpy_b200 import synth
    det, threads = cpu_reference_detector(a.width, a.height, a.window, a.fps, dy=not a.no_dy)
    # a step = a bounded sample of the workload: 8 frames, fewer when many steps are asked for, so that the whole
    # run (frame synthesis on the host included) stays within a few minutes
    per_step = max(1, min(8, 240 // max(1, a.warmup + a.steps)))
    sky = synth.make_sky(a.width, a.height)
    total = a.window + 2 + (a.warmup + a.steps) * per_step
    frames = [synth.make_frame(t, sky, a.width, a.height, a.fps) for t in range(total)]
    t = 0
    for _ in range(a.window + 2):  # fill the window
        det.update(frames[t]); det.detect(); t += 1
    for _ in range(a.warmup):
        for _ in range(per_step):
            det.update(frames[t]); det.detect(); t += 1
    t0 = time.perf_counter()
    for _ in range(a.steps):
        for _ in range(per_step):
            det.update(frames[t]); det.detect(); t += 1
    dt = time.perf_counter() - t0
    fps = a.steps * per_step / dt
    sample = f"{per_step} frames/step after a {a.window + 2}-frame window fill; numpy single-thread + cv2 {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a)}, "frames_per_step": per_step,
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def bind_near_gpu(local):
    """N > 1: pin this rank's threads to the CPUs of its GPU's NUMA node (NVML's ideal affinity), so that the pinned
    staging buffers it allocates next are first-touched on that node and every rank's host->device copies stay behind
    their own root complex instead of all ranks sharing the cores and memory of node 0.  Returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        before = len(os.sched_getaffinity(0))
        words = (os.cpu_count() + 63) // 64
        maskw = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = {64 * i + b for i, w in enumerate(maskw) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        node = None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(hnd).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        except Exception:
            pass
        if cpus and len(cpus) < before:
            os.sched_setaffinity(0, cpus)
            return {"bound": True, "cpus": len(cpus), "of": before, "numa_node": node}
        return {"bound": False, "cpus": before, "numa_node": node, "why": "NVML reports no narrower CPU set for this GPU"}
    except Exception as e:
        return {"bound": False, "why": repr(e)}


# ---------------------------------------------------------------------------------------------
class LineGather:
    """All-gather of line records (frame, x1, y1, x2, y2, nonline_prob), one call per step, on a side stream with the
    collective launched asynchronously: nothing on the submitting thread waits for it before the end of the job."""
    CAP = 16384
    SLOTS = 4

    def __init__(self, world, dev):
        import torch
        self.world, self.dev = world, dev
        self.stream = torch.cuda.Stream(device=dev, priority=-1)
        self.host = [torch.zeros((self.CAP, 6), dtype=torch.float64).pin_memory() for _ in range(self.SLOTS)]
        self.devb = [torch.zeros((self.CAP, 6), dtype=torch.float64, device=dev) for _ in range(self.SLOTS)]
        self.all = [torch.zeros((world * self.CAP, 6), dtype=torch.float64, device=dev) for _ in range(self.SLOTS)]
        self.work = [None] * self.SLOTS
        self.copied = [None] * self.SLOTS  # event: the pinned buffer of the slot has been read by its H2D copy
        self.k = 0
        self.fill = 0
        self.total = torch.zeros((), dtype=torch.float64, device=dev)
        self.dropped = 0

    def add_batch(self, det, T, first_frame):
        from metdetpy_b200 import sharding as S
        s = self.k % self.SLOTS
        if self.fill == 0 and self.copied[s] is not None:
            self.copied[s].synchronize()  # four steps old: long done
        rows = self.host[s].numpy()
        room = self.CAP - 1 - self.fill
        k = S.pack_line_records(det, T, first_frame, rows[self.fill:self.fill + room])
        want = int(det.last_infos["n_lines"][:T].sum())
        self.dropped += want - k
        self.fill += k

    def flush(self):
        """end of a step: ship what has been packed"""
        import torch
        import torch.distributed as dist
        s = self.k % self.SLOTS
        self.host[s].numpy()[self.CAP - 1, 0] = self.fill
        with torch.cuda.stream(self.stream):
            self.devb[s].copy_(self.host[s], non_blocking=True)
            self.copied[s] = torch.cuda.Event()
            self.copied[s].record(self.stream)
            self.work[s] = dist.all_gather_into_tensor(self.all[s], self.devb[s], async_op=True)
            self.work[s].wait()  # orders the side stream after the collective; the host does not block
            self.total += self.all[s].view(self.world, self.CAP, 6)[:, self.CAP - 1, 0].sum()
        self.k += 1
        self.fill = 0

    def finish(self):
        """waits for the outstanding exchanges; returns the number of records all ranks shipped since the last finish"""
        self.stream.synchronize()
        v = int(self.total.item())
        self.total.zero_()
        return v


def main_ours(a):
    import torch
    import torch.distributed as dist
    from metdetpy_b200 import sharding as S
    from metdetpy_b200.detector import M3Detector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(local) if world > 1 else None  # before any pinned allocation (first touch decides the node)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    W, H, n, B, bps = a.width, a.height, a.window, a.batch, a.bps
    HW = W * H
    if world > 1 and bps % DISTINCT:
        raise SystemExit(f"--bps must be a multiple of {DISTINCT} for --gpus > 1")
    mask = np.ones((H, W), np.uint8)
    det = M3Detector(n / a.fps + 1e-9, a.fps, mask, 10, make_cfg(dy=not a.no_dy), None, device=local, max_batch=B)
    if a.generic_kernel:
        det._eng.set_option("stream_kernel", 0)
    for kv in os.environ.get("MDB_OPTS", "").split(","):
        if "=" in kv:
            det._eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    ext = torch.cuda.ExternalStream(det._eng.stream_ptr(), device=dev)
    pad = 2 * n - 2
    stream = Stream(B, W, H, a.fps, dev, pad)
    _VIEWS["main"] = stream
    interval = 2
    roi = det.stack.std_roi
    roi_px = (roi[2] - roi[0]) * (roi[3] - roi[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        el = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        return float(el.item())

    trace = [] if os.environ.get("BENCH_TRACE") else None
    acc = {"fused_ms": 0.0, "batches": 0, "lines": 0}

    # ---- N = 1: one sequential stream, three batches in flight ---------------------------------------------------
    def sequential(first_batch, nbatches, on_batch=None, timed=False):
        AHEAD = 2
        for s in range(min(AHEAD, nbatches)):
            det.submit(stream.ptr((first_batch + s) * B), B, True)
        for s in range(nbatches):
            if s + AHEAD < nbatches:
                det.submit(stream.ptr((first_batch + s + AHEAD) * B), B, True)
            res = det.collect()
            if timed:
                acc["lines"] += sum(len(r[0]) for r in res)
                acc["fused_ms"] += det._eng.fused_time()[0]
                acc["batches"] += 1
                if trace is not None and (s + 1) % bps == 0:
                    trace.append(time.perf_counter())
            if on_batch is not None:
                on_batch(det, (first_batch + s) * B)

    # ---- N > 1: one job of the time-sharded algorithm over `steps` steps per rank --------------------------------
    def sharded_job(steps, vrank=rank, vworld=world, gather=None, digests=None, timed=False, group_exchange=True):
        C = steps * bps * B                       # frames per rank
        sh = S.Shard(vrank, vrank * C, (vrank + 1) * C, max(0, vrank * C - pad))
        tp = [time.perf_counter()]
        # phase A: integer noise sums of the sample timers of this rank's chunk
        segs = [S.Segment(stream.ptr(t), B, t, history=pad) for t in range(sh.start, sh.end, B)]
        mine = S.chunk_noise_samples(det, segs, n, interval, sh.start, sh.end, HW)
        tp.append(time.perf_counter())
        # phase B: all-gather of the triples (one packed int64 tensor; every rank owns the same number of samples +-1)
        if group_exchange and world > 1:
            cap = C // (interval * n) + n + 2
            buf = torch.zeros((cap, 3), dtype=torch.int64)
            buf[0, 0] = len(mine)
            buf[1:1 + len(mine)] = torch.from_numpy(mine)
            allb = torch.zeros((world * cap, 3), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allb, buf.to(dev))
            allh = allb.cpu().view(world, cap, 3).numpy()
            # later ranks' samples are not needed for this rank's thresholds
            samples = np.concatenate([allh[r, 1:1 + int(allh[r, 0, 0])] for r in range(vrank + 1)])
        else:
            parts = []
            for r in range(vrank):      # virtual ranks on one GPU: compute the earlier chunks' samples here
                segs_r = [S.Segment(stream.ptr(t), B, t, history=pad) for t in range(r * C, (r + 1) * C, B)]
                parts.append(S.chunk_noise_samples(det, segs_r, n, interval, r * C, (r + 1) * C, HW))
            samples = np.concatenate(parts + [mine])
        tp.append(time.perf_counter())
        # phase C: bit-identical threshold replay for the frames this rank ingests
        thr, thr_f, snr = S.replay_thresholds_native(samples, roi_px, n, sh.halo_start, sh.end, adaptive=True, init_value=7,
                                                     sensitivity="normal", interval=interval)
        tp.append(time.perf_counter())
        # phase D: seek + halo batch + chunk, three batches in flight; line records shipped once per step
        run = ([S.Segment(stream.ptr(sh.halo_start), sh.start - sh.halo_start, sh.halo_start)] if sh.start > sh.halo_start else []) + segs
        state = {"k": 0}

        def on_batch(d, sg):
            state["k"] += 1
            if timed:
                acc["fused_ms"] += d._eng.fused_time()[0]
                acc["batches"] += 1
            if digests is not None:
                digests.append(S.batch_digest(d, sg.T))
            if gather is not None:
                gather.add_batch(d, sg.T, sg.t0)
                if state["k"] % bps == 0:
                    gather.flush()
            elif timed:
                acc["lines"] += int(d.last_infos["n_lines"][:sg.T].sum())
            if trace is not None and timed and state["k"] % bps == 0:
                trace.append(time.perf_counter())
        S.run_chunk(det, run, sh, thr, thr_f, snr, thr_base=sh.halo_start, on_batch=on_batch)
        tp.append(time.perf_counter())
        if trace is not None and timed:
            d = np.diff(tp) * 1e3
            print(f"rank {rank} phases (ms): noise sums {d[0]:.2f}, all-gather {d[1]:.2f}, replay {d[2]:.2f}, "
                  f"seek + halo + chunk {d[3]:.2f} ({len(mine)} own samples, {len(samples)} replayed)", file=sys.stderr)
        return sh

    # ---- warm-up + timed region ------------------------------------------------------------------------------------
    gather = LineGather(world, dev) if world > 1 else None
    digests = []
    if world == 1:
        sequential(0, a.warmup * bps)
    else:
        sharded_job(max(1, a.warmup), gather=gather)
        gather.finish()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = det._eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(ext)
    if world == 1:
        sequential(a.warmup * bps, a.steps * bps, timed=True)
        nlines_total = acc["lines"]
    else:
        sharded_job(a.steps, gather=gather, digests=digests, timed=True)
        nlines_total = gather.finish()
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop()
    launches = det._eng.launch_count() - l0
    tiers = {k: int(det._eng.info(k + "_total")) for k in ("hough_tier1a", "hough_tier1b", "hough_tier2", "hough_tier3")}
    wall_max = allmax(wall)
    value = world * a.steps * bps * B / wall_max
    dev_ms = e0.elapsed_time(e1)
    fused_ms, fused_batches = acc["fused_ms"], acc["batches"]
    if trace:
        steps_ms = np.diff(np.array([t0] + trace)) * 1e3
        print(f"rank {rank} step times (ms): " + " ".join(f"{x:.2f}" for x in steps_ms), file=sys.stderr)

    # ---- parity of the time-sharded result against one sequential pass of the same frames ------------------------
    parity = None
    if not a.no_parity:
        if world > 1:
            nb = (rank + 1) * a.steps * bps
            ok, scope = 1, "every rank re-ran frames 0 .. end of its chunk sequentially"
            if nb <= a.parity_budget:
                det.reset()
                seq = []
                first = rank * a.steps * bps

                def on_b(d, t0_):
                    if t0_ >= first * B:
                        seq.append(S.batch_digest(d, B))
                sequential(0, nb, on_batch=on_b)
                ok = int(seq == digests and len(seq) == a.steps * bps)
                checked = 1
            else:
                checked = 0
            okt = torch.tensor([ok, checked], device=dev, dtype=torch.int64)
            mn = okt.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            sm = okt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            parity = {"ok": bool(mn[0].item()), "ranks_checked": int(sm[1].item()), "frames_per_rank": a.steps * bps * B,
                      "scope": scope if int(sm[1].item()) == world else
                      f"ranks whose sequential prefix fits {a.parity_budget} batches re-ran frames 0 .. end of their chunk",
                      "compared": "per-batch digests of bi_threshold, on-pixel count, raw segment count and raw Hough segments of every frame"}
        else:
            # one GPU: the sharded protocol with two virtual ranks over a short stream, against the sequential pass
            st = 1
            seq = []
            det.reset()
            sequential(0, 2 * st * bps, on_batch=lambda d, t0_: seq.append(S.batch_digest(d, B)))
            got = []
            for vr in range(2):
                dg = []
                sharded_job(st, vrank=vr, vworld=2, digests=dg, group_exchange=False)
                got += dg
            parity = {"ok": bool(got == seq and len(seq) == 2 * st * bps), "ranks_checked": 2, "frames_per_rank": st * bps * B,
                      "scope": "two virtual ranks on this GPU (reset + seek + halo + replayed thresholds) against one sequential pass",
                      "compared": "per-batch digests of bi_threshold, on-pixel count, raw segment count and raw Hough segments of every frame"}
        det.reset()
    barrier()

    # ---- the mask chain timed alone (roofline): one batch in flight, so nothing else shares the SMs;
    # CUDA events on the library's own streams bracket the temporal pass (front stream) and act + dst (back stream)
    det.reset()
    alone_ms, alone_t, alone_launches, alone_steps = 0.0, 0.0, 0, 8
    for s in range(alone_steps + 2):
        det.submit(stream.ptr(s * B), B, True)
        det.collect(want_lines=False)
        if s >= 2:
            ms, nl = det._eng.fused_time()
            alone_ms += ms; alone_launches += nl
            alone_t += det._eng.info("temporal_ms")
    temporal_gen = int(det._eng.info("temporal_generation"))
    barrier()

    # ---- end to end through the public API with HOST (pinned) buffers -------------------------
    e2e = None
    if not a.no_e2e:
        hosts, ok = [], 1
        try:  # two pinned staging buffers per rank (B*H*W bytes each); all ranks must agree to go on
            for s in range(2):
                hb = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
                hb.copy_(stream.view(s * B, B))
                hosts.append(hb)
        except Exception as exc:
            print(f"rank {rank}: no pinned host memory for the end-to-end leg: {exc!r}", file=sys.stderr)
            hosts, ok = [], 0
        okt = torch.tensor([ok], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        torch.cuda.synchronize()
        if int(okt.item()) == 1:
            det.reset()
            nb = max(8, min(a.steps * bps, 24))

            def submit_host(s):
                det.submit(hosts[s % len(hosts)].data_ptr(), B, False)
            for s in range(3):
                submit_host(s)
                det.collect()
            barrier()
            t0e = time.perf_counter()
            submit_host(0)
            for s in range(nb):
                if s + 1 < nb:
                    submit_host(s + 1)  # its H2D copy overlaps the kernels of batch s
                det.collect()           # D2H of the results + host NMS
            barrier()
            w2 = allmax(time.perf_counter() - t0e)
            d2h = B * (4 + 8 + 8 + 4 + 4 + 512 * 16)  # thr, thr_float, snr, n_on, n_lines, raw segments
            e2e = {"value": world * nb * B / w2, "unit": UNIT, "h2d_bytes_per_step": bps * B * HW,
                   "d2h_bytes_per_step": bps * d2h, "batches_timed": nb,
                   "h2d_gbs_per_rank": nb * B * HW / w2 / 1e9, "numa_binding_rank0": numa,
                   "mode": "one host-fed stream per rank (pinned buffers -> staging copy -> zero-copy kernels), H2D of batch k+1 "
                           "overlapping the kernels of batch k"}
            del hosts

    extra = {}
    if rank == 0 and world == 1:
        peaks = _peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        if not a.no_extra:
            for name, fn in (("per_frame_api", lambda: measure_per_frame(a, det, stream, dev)),
                             ("dense_regime", lambda: measure_dense(a, stream, dev))):
                try:
                    extra[name] = fn()
                except Exception as e:  # a side measurement must never take the headline line down
                    extra[name] = {"error": repr(e)}
        if not a.no_next_rows:
            try:
                extra["next_rows"] = measure_next_rows(a, stream.ptr(0), dev, peak)
            except Exception as e:
                extra["next_rows"] = {"error": repr(e)}
    det.close()
    if rank == 0 and world == 1 and not a.no_configs:
        del stream.buf
        _VIEWS.clear()
        torch.cuda.empty_cache()
        try:
            extra["configs"] = measure_configs(a, dev, float(_peaks().get("hbm_gbs", 6650.0)))
        except Exception as e:
            extra["configs"] = {"error": repr(e)}

    if rank == 0:
        peaks = _peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = 2.0 * HW * B * alone_steps  # read the new u8 frame once + write the u8 mask once
        achieved = alg_bytes / (alone_ms * 1e-3) / 1e9 if alone_ms > 0 else None
        in_step = 2.0 * HW * B * fused_batches / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": wall_max / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(a)},
            "run": {"frames_per_step_per_gpu": bps * B, "batches_per_step": bps, "frames_per_call": B,
                    "l2_policy": f"inputs larger than L2 ({B * HW / 1e6:.0f} MB per batch, {DISTINCT} batches cycled)",
                    "sharding": ("one stream of N time chunks: (2n-2)-frame halo, all-gather of integer noise sums, threshold replay, "
                                 "line records all-gathered once per step on a side stream; no frame crosses GPUs") if world > 1 else
                                "one sequential stream, three batches in flight"},
            "sharded_parity": (parity or {}).get("ok"), "parity": parity,
            "device_ms_per_step": dev_ms / a.steps,
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": _traffic_from_profile(W, H, n, B),
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of the chain) "
                                         "from the committed ncu --set full capture of this configuration under profiles/ (null: none committed)",
                         "algorithmic_bytes_per_launch": 2.0 * HW * B * alone_steps / max(alone_launches, 1),
                         "kernel": f"fused mask chain: temporal{temporal_gen}_kernel (stack->diff->threshold) + act4_kernel (median+close) + dst_sparse/dense_kernel (dy-mask, mask bytes)",
                         "kernel_ms_per_launch": alone_ms / max(alone_launches, 1),
                         "kernel_launches": alone_launches,
                         "chain_ms_per_batch": alone_ms / alone_steps,
                         "temporal_ms_per_batch": alone_t / alone_steps,
                         "timing": f"CUDA events on the library's streams around the chain, {alone_steps} batches with one batch in "
                                   "flight (chain alone on the GPU) right after the timed region",
                         "achieved_inside_timed_region": in_step,
                         "frac_inside_timed_region": (in_step / peak) if in_step else None,
                         "inside_note": "same events during the timed region, where the chain shares the SMs with the Hough pass "
                                        "and the next batch's temporal pass (three batches in flight): elapsed, not busy, time",
                         "frac_whole_step": 2.0 * HW * B * bps / (dev_ms / a.steps * 1e-3) / 1e9 / peak if dev_ms else None,
                         "whole_step_note": "algorithmic bytes of a step / the device time of the step (every kernel of the path: noise, "
                                            "thresholds, mask chain, PPHT, result copies; three batches in flight) against the same peak",
                         "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s"},
            "clocks": sampler.summary(), "nms_lines_total": nlines_total,
            "ppht_tiers_rank0": tiers,
        }
        if gather is not None and gather.dropped:
            out["line_records_dropped"] = gather.dropped
        out.update(extra)
        if not a.no_cpu_baseline and world == 1:
            v, threads, warm = run_cpu_sample(a, a.cpu_frames)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{a.cpu_frames} frames of the same stream after a {warm}-frame window fill; "
                                             f"oracle cv2 backend = the reference's numpy+cv2 calls; cv2 threads {threads}, numpy 1"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _traffic_from_profile(W, H, n, B):
    """DRAM bytes per launch of the mask chain from the committed ncu capture of this configuration
    (profiles/chain_traffic.json, written by profiles/ncu_summary.py --traffic from an ncu --set full raw CSV)."""
    try:
        tab = json.load(open(os.path.join(REPO, "profiles", "chain_traffic.json")))
        e = tab.get(f"{W}x{H}_n{n}_b{B}")
        return e["bytes_per_launch"] if e else None
    except Exception:
        return None


def measure_per_frame(a, det, stream, dev):
    """The reference's own call pattern (MetDetPy.py:197-198): update(frame); detect() per frame, frames in pinned host
    memory.  Two streams: the bench stream (a meteor trail in almost every frame: the sequential exact-order PPHT of
    a ~1000-point trail is most of a frame's time) and a quiet sky (noise only -- what nearly every frame of a real
    night is: the figure is then PCIe + launch bound)."""
    import torch
    W, H = a.width, a.height
    F = 96
    host = torch.empty((F, H, W), dtype=torch.uint8).pin_memory()
    out = {"unit": UNIT, "frames": F - 32,
           "call": "M3Detector.update(frame); M3Detector.detect() per frame from pinned host memory (MetDetPy.py:197-198)",
           "h2d_bytes_per_frame": W * H,
           "path": "resident window state, O(1) in the window length (csrc/perframe_kernel.cuh); update() is asynchronous"}
    for key in ("value", "quiet_sky"):
        if key == "value":
            host.copy_(stream.view(0, F))
        else:
            g = torch.Generator(device=dev); g.manual_seed(7)
            sky = (torch.randn((F, H, W), device=dev, generator=g) * 2.0 + 40.0).clamp_(0, 255).to(torch.uint8)
            host.copy_(sky)
            del sky
        frames = host.numpy()
        det.reset()
        for t in range(32):
            det.update(frames[t]); det.detect()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(32, F):
            det.update(frames[t]); det.detect()
        dt = time.perf_counter() - t0
        out[key] = (F - 32) / dt
    out["temporal_generation"] = int(det._eng.info("temporal_generation"))
    return out


def measure_dense(a, stream, dev):
    """Dense-mask regime (~5e4 on-pixels per frame on the noisy real 4K clip): the same 4K stream with a
    fixed low threshold so that sensor noise lights up tens of thousands of pixels per frame; every frame takes the
    global-memory PPHT tiers.  CPU arm: the reference's numpy + cv2 calls on the same frames."""
    import torch
    from metdetpy_b200.detector import M3Detector
    W, H, n = a.width, a.height, a.window
    T = 128
    mask = np.ones((H, W), np.uint8)
    out = None
    for thr in (5,):
        det = M3Detector(n / a.fps + 1e-9, a.fps, mask, 10, make_cfg(dy=True, adaptive=False, init_value=thr), None,
                         device=dev.index or 0, max_batch=T)
        det.detect_many((stream.ptr(0), T), on_device=True)          # fills the window
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        det.detect_many((stream.ptr(T), T), on_device=True)
        dt = time.perf_counter() - t0
        non = det.last_infos["n_on"].astype(np.float64)
        raw = det.last_infos["lines_num"].astype(np.float64)
        det_info = {k: det._eng.info(k) for k in ("hough_tier1a", "hough_tier1b", "hough_tier2", "hough_tier3")}
        det.close()
        out = {"threshold": thr, "frames": T, "value": T / dt, "unit": UNIT, "on_pixels_mean": float(non.mean()),
               "on_pixels_max": float(non.max()), "raw_segments_mean": float(raw.mean()),
               "ppht_tiers": {k: int(det_info[k]) for k in ("hough_tier1a", "hough_tier1b", "hough_tier2", "hough_tier3")}}
        if non.mean() >= 5000:
            break
    # CPU arm on a few of the same frames
    ref, threads = cpu_reference_detector(W, H, n, a.fps, dy=True, adaptive=False, init_value=out["threshold"])
    host = stream.view(T - n - 2, n + 2 + 6).cpu().numpy()
    for f in host[:n + 2]:
        ref.update(f); ref.detect()
    t0 = time.perf_counter()
    for f in host[n + 2:]:
        ref.update(f); ref.detect()
    cpu = 6 / (time.perf_counter() - t0)
    out["cpu_baseline"] = {"value": cpu, "unit": UNIT, "kind": "port", "cores": os.cpu_count(),
                           "sample": f"6 frames of the same stream, fixed threshold {out['threshold']}; cv2 threads {threads}"}
    out["speedup_vs_cpu"] = out["value"] / cpu
    return out


def measure_configs(a, dev, peak):
    """The other BASELINE.json configs on ONE GPU (device-resident frames, three batches in flight): frames/s, the mask
    chain alone and its fraction of the HBM roofline."""
    import torch
    from metdetpy_b200.detector import M3Detector
    res = {}
    cases = [("config2_1080p_n5", 1920, 1080, 30.0, 5, False, None, 2048),
             ("config4_4k60_n60_mask", 3840, 2160, 60.0, 60, True, "mask-east", 512),
             ("config5_8k_n30", 7680, 4320, 30.0, 30, True, None, 256)]
    for name, W, H, fps, n, dy, mk, B in cases:
        HW = W * H
        mask = np.ones((H, W), np.uint8)
        note = None
        if mk:
            mask, note = load_bench_mask(W, H)
        det = M3Detector(n / fps + 1e-9, fps, mask, 10, make_cfg(dy=dy), None, device=dev.index or 0, max_batch=B,
                         apply_mask=mk is not None)
        st = Stream(B, W, H, fps, dev, 0, distinct=2)
        nb = 10
        for s in range(3):
            det.submit(st.ptr(s * B), B, True); det.collect()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        det.submit(st.ptr(3 * B), B, True); det.submit(st.ptr(4 * B), B, True)
        for s in range(nb):
            if s + 2 < nb:
                det.submit(st.ptr((5 + s) * B), B, True)
            det.collect()
        dt = time.perf_counter() - t0
        chain = []
        for s in range(5):
            det.submit(st.ptr(s * B), B, True); det.collect(want_lines=False)
            chain.append(det._eng.fused_time()[0])
        chain_ms = float(np.median(chain[1:]))
        res[name] = {"workload": f"synthetic {W}x{H} @{fps:g}fps, window={n}, dy_mask={'on' if dy else 'off'}, "
                                 f"{'mask applied on the device, ' if mk else ''}{B} frames per call",
                     "frames_per_s": nb * B / dt, "chain_ms_per_call": chain_ms,
                     "temporal_generation": int(det._eng.info("temporal_generation")),
                     "roofline_frac": 2.0 * HW * B / (chain_ms * 1e-3) / 1e9 / peak}
        if note:
            res[name]["mask"] = note
        det.close()
        del st
        torch.cuda.empty_cache()
    return res


def load_bench_mask(W, H):
    """Config 4's mask: the reference's bundled test/mask-east.jpg through fileio.load_mask at this size, committed
    bit-packed under tests/golden/ (made by tests/golden/make_golden.py); a synthetic horizon mask if it is absent."""
    p = os.path.join(REPO, "tests", "golden", f"mask_east_{W}x{H}.npz")
    if os.path.exists(p):
        z = np.load(p)
        m = np.unpackbits(z["bits"])[:H * W].reshape(H, W).astype(np.uint8)
        return m, "test/mask-east.jpg through fileio.load_mask (tests/golden)"
    m = np.ones((H, W), np.uint8)
    yy = np.arange(H)[:, None]
    xx = np.arange(W)[None, :]
    m[yy > H * 0.82 + 0.05 * H * np.sin(xx / W * 7.0)] = 0
    return m, "synthetic horizon mask (tests/golden/mask_east file absent)"


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
