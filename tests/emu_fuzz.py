"""Randomised parity runs WITHOUT a GPU: the product's kernels under the CPU block emulator (tests/emu/stream_path_emu.cpp:
the batched streaming path and the per-frame resident-state path) against the CPU checker (oracle/m3_oracle.py, cv2 backend =
the reference's own call sites) on random small configurations -- frame sizes, every window that has a temporal3 shape, batch
lengths that cut the van Herk blocks anywhere, adaptive / fixed thresholds, dynamic mask on / off, Hough parameters, masks,
bright flashes and stuck hot regions.  Test tooling (tests/test_emu_fuzz_cpu.py runs a few seeds; `python tests/emu_fuzz.py
FIRST COUNT` runs more)."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

T3_WINDOWS = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20, 21, 24, 25, 28, 30, 32, 36, 40, 48, 50, 60, 64]
_SENS = ["low", "normal", "high"]


def build_lib(tmp):
    from emu_build import build_stream
    lib = C.CDLL(build_stream(tmp))
    lib.emu_stream_path.restype = C.c_int
    lib.emu_perframe_path.restype = C.c_int
    lib.emu_temporal2_launches.restype = C.c_int
    return lib


def build_generic_lib(tmp):
    from emu_build import build_generic
    lib = C.CDLL(build_generic(tmp))
    lib.emu_generic_path.restype = C.c_int
    return lib


def make_case(seed, any_width=False, dense=False, big=False):
    import cv2
    r = np.random.default_rng(seed)
    W = int(r.integers(17, 260)) if any_width else 32 * int(r.integers(1, 9))
    H = int(r.choice([8, 9, 15, 16, 17, 31, 33, 40, 63, 64, 65, 90]))
    if big:  # wide frames: many 32-px words per row (strips of the act kernel, 16-byte chunks of act4), several row bands
        W = 32 * int(r.choice([16, 32, 41, 60])) + (int(r.integers(1, 32)) if any_width else 0)
        H = int(r.choice([65, 130, 200]))
    n = int(r.choice(T3_WINDOWS)) if r.random() < 0.8 else int(r.integers(2, 65))  # windows without a shape are skipped by the caller
    if big:
        n = int(r.choice([2, 5, 6, 12, 17, 30]))
    if any_width and r.random() < 0.15:
        n = int(r.choice([1, 70, 129, 140]))  # the generic kernels take any window
    T = int(min(110, r.integers(max(2, n // 2), 2 * n + 24)))
    if big:
        T = int(min(T, n + 8))
    batch = int(r.integers(1, T + 1)) if r.random() < 0.7 else int(r.integers(1, 9))
    cfg = dict(adaptive=bool(r.random() < 0.6), init_value=int(r.integers(3, 13)), sensitivity=_SENS[int(r.integers(0, 3))],
               area=float(r.choice([0.1, 0.2, 0.4])), interval=int(r.integers(1, 4)),
               hough=(int(r.integers(4, 16)), int(r.integers(3, 16)), int(r.integers(0, 11))), dy_mask=bool(r.random() < 0.6))
    mask = np.ones((H, W), np.uint8)
    if r.random() < 0.5:  # a polygon cut out of a corner, like a horizon mask
        pts = np.array([[0, H], [0, int(H * r.uniform(0.5, 0.95))], [int(W * r.uniform(0.3, 1.0)), H]], np.int32)
        cv2.fillPoly(mask, [pts], 0)
    sigma = float(r.uniform(0.8, 3.5))
    if dense:  # low fixed threshold on a noisy sky: thousands of on-pixels per frame (PPHT tiers 1b / 2 / 3, dst_dense, list overflows)
        cfg.update(adaptive=False, init_value=int(r.integers(1, 4)))
        sigma = float(r.uniform(3.0, 6.0))
        T = min(T, 24)
    sky = r.uniform(10, 60) + r.normal(0, 3, (H, W))
    frames = np.empty((T, H, W), np.uint8)
    ang, x0, y0 = r.uniform(0, 2 * np.pi), r.uniform(0.1, 0.9) * W, r.uniform(0.1, 0.9) * H
    speed, bright, thick = r.uniform(1.0, 6.0), r.uniform(25, 120), int(r.integers(1, 4))
    start = int(r.integers(0, max(1, T - 3)))
    hot = (int(r.integers(0, H - 3)), int(r.integers(0, W - 6))) if r.random() < 0.4 else None
    flash = int(r.integers(0, T)) if r.random() < 0.3 else -1
    for t in range(T):
        f = sky + r.normal(0, sigma, (H, W))
        if t >= start:
            k = t - start
            p1 = (int(x0 + speed * k * np.cos(ang)), int(y0 + speed * k * np.sin(ang)))
            p2 = (int(x0 + speed * (k + 2) * np.cos(ang)), int(y0 + speed * (k + 2) * np.sin(ang)))
            m = np.zeros((H, W), np.float32)
            cv2.line(m, p1, p2, float(bright), thick, cv2.LINE_AA)
            f = f + m
        if hot is not None and t % 3 != 2:  # a flickering hot region: what the dynamic mask is for
            f[hot[0]:hot[0] + 3, hot[1]:hot[1] + 6] += 90
        if t == flash:
            f = f + r.uniform(20, 80)
        frames[t] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    raw_frames = frames.copy()
    frames *= mask  # what the loader hands to the detector (MetLib/imgproc.py:96-101)
    return dict(W=W, H=H, n=n, T=T, batch=batch, cfg=cfg, mask=mask, frames=frames, raw_frames=raw_frames,
                apply_mask=bool(r.random() < 0.5))


def run_case(lib, case, per_frame=False, generic=False):
    """None if equal, else a description of the first difference.  generic: `lib` is build_generic_lib()'s -- the per-frame
    generic kernels (any width, any window), optionally with the mask applied inside the device loads."""
    from metdetpy_b200.detector import select_subarea
    from oracle import m3_oracle as O
    W, H, n, T, cfg, mask, fr = (case[k] for k in ("W", "H", "n", "T", "cfg", "mask", "frames"))
    fps = 25.0
    roi_t = select_subarea(mask, cfg["area"])
    roi = (C.c_int * 4)(*[int(v) for v in roi_t])
    thr = np.zeros(T, np.int32); snr = np.zeros(T)
    dst = np.zeros((T, H, W), np.uint8); n_on = np.zeros(T, np.int32); nl = np.zeros(T, np.int32)
    raw = np.zeros((T, 512, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    common = (int(cfg["adaptive"]), int(cfg["init_value"]), _SENS.index(cfg["sensitivity"]), int(cfg["interval"]), roi,
              *[int(v) for v in cfg["hough"]], int(cfg["dy_mask"]), C.c_double(float(np.sum(mask))),
              p(thr), p(snr), p(dst), p(n_on), p(nl), p(raw))
    dev_mask = bool(case["apply_mask"]) and not generic and hasattr(lib, "emu_set_device_mask")
    if dev_mask:  # apply_mask = 1: unmasked frames in, the mask is applied inside the kernels' loads
        fr = case["raw_frames"]
        lib.emu_set_device_mask(p(mask))
    if generic:
        thrf = np.zeros(T)
        src = case["raw_frames"] if case["apply_mask"] else fr
        rc = lib.emu_generic_path(p(src), T, W, H, n, p(mask), int(case["apply_mask"]), *common[:-6], p(thr), p(thrf), *common[-5:])
    elif per_frame:
        rc = lib.emu_perframe_path(p(fr), T, W, H, n, *common)
    else:
        rc = lib.emu_stream_path(p(fr), T, W, H, n, case["batch"], *common)
    if dev_mask:
        lib.emu_set_device_mask(None)
        fr = case["frames"]
    if rc in (-1000, -1001):  # width not a multiple of 32 / window without a temporal3 shape (temporal2 on the device)
        return "skipped"
    if rc != 0:
        return f"emulated path failed: rc={rc}"
    ref = O.M3DetectorOracle(n / fps + 1e-9, fps, mask, 10, adaptive=cfg["adaptive"], init_value=cfg["init_value"],
                             sensitivity=cfg["sensitivity"], area=cfg["area"], interval=cfg["interval"], hough=cfg["hough"],
                             dy_mask=cfg["dy_mask"], backend="cv2")
    assert ref.stack_maxsize == n and tuple(ref.stack.std_roi) == tuple(roi_t)
    for t in range(T):
        ref.update(fr[t]); ref.detect()
        if ref.bi_threshold != thr[t]:
            return f"frame {t}: threshold {thr[t]} != {ref.bi_threshold}"
        if not np.isclose(snr[t], ref.stack.snr, rtol=1e-12, atol=0):
            return f"frame {t}: snr {snr[t]!r} != {ref.stack.snr!r}"
        if not np.array_equal(dst[t], ref.dst):
            return f"frame {t}: mask differs in {int(np.count_nonzero(dst[t] != ref.dst))} pixels"
        if nl[t] != ref.lines_num:
            return f"frame {t}: {nl[t]} raw segments != {ref.lines_num}"
        if 0 < nl[t] <= 500 and not np.array_equal(raw[t, :nl[t]], np.asarray(ref.linesp_ext).reshape(-1, 4)):
            return f"frame {t}: raw segments differ"
    return None


def run_sharded_case(lib, case, world):
    """SURVEY 8(e) on a random case: per-rank noise sums -> pooled -> mdb_replay_thresholds -> seek + halo + chunk per virtual rank
    (all emulated product kernels + the library's host code) against the sequential emulated run of the same case."""
    from metdetpy_b200 import sharding as S
    from metdetpy_b200.detector import select_subarea
    W, H, n, T, cfg, mask, fr = (case[k] for k in ("W", "H", "n", "T", "cfg", "mask", "frames"))
    if T < world or n > 128:
        return "skipped"
    roi_t = [int(v) for v in select_subarea(mask, cfg["area"])]
    roi = (C.c_int * 4)(*roi_t)
    roi_px = (roi_t[2] - roi_t[0]) * (roi_t[3] - roi_t[1])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hough = [int(v) for v in cfg["hough"]]
    area = C.c_double(float(np.sum(mask)))
    # sequential run
    s_thr = np.zeros(T, np.int32); s_snr = np.zeros(T); s_dst = np.zeros((T, H, W), np.uint8)
    s_on = np.zeros(T, np.int32); s_nl = np.zeros(T, np.int32); s_raw = np.zeros((T, 512, 4), np.int32)
    rc = lib.emu_stream_path(p(fr), T, W, H, n, case["batch"], int(cfg["adaptive"]), int(cfg["init_value"]), _SENS.index(cfg["sensitivity"]),
                             int(cfg["interval"]), roi, *hough, int(cfg["dy_mask"]), area, p(s_thr), p(s_snr), p(s_dst), p(s_on), p(s_nl), p(s_raw))
    if rc != 0:
        return "skipped" if rc in (-1000, -1001) else f"sequential run failed: rc={rc}"
    shards = S.plan_shards(T, world, n)
    samples = []
    for sh in shards:
        taus = np.ascontiguousarray(S.sample_timers(sh.start, sh.end, n, int(cfg["interval"])), np.int64)
        if len(taus) == 0:
            continue
        lo = max(0, sh.start - (n - 1))
        part = np.ascontiguousarray(fr[lo:sh.end])
        sums = np.zeros((len(taus), 2), np.uint64)
        if lib.emu_noise_sums(p(part), len(part), C.c_longlong(lo), W, H, n, int(cfg["interval"]), roi, p(taus), len(taus), p(sums)) != 0:
            return f"rank {sh.rank}: emu_noise_sums failed"
        samples += [(int(t), int(a), int(b)) for t, (a, b) in zip(taus, sums)]
    thr, thr_f, snr = S.replay_thresholds_native(samples, roi_px, n, 0, T, adaptive=cfg["adaptive"], init_value=cfg["init_value"],
                                                 sensitivity=cfg["sensitivity"], interval=cfg["interval"])
    if not np.array_equal(thr, s_thr):
        return f"replayed thresholds differ at frame {int(np.flatnonzero(np.asarray(thr) != s_thr)[0])}"
    if not np.allclose(snr, s_snr, rtol=1e-12, atol=0):
        return "replayed snr differs"
    for sh in shards:
        part = np.ascontiguousarray(fr[sh.halo_start:sh.end])
        m, halo = len(part), sh.start - sh.halo_start
        thr_in = np.ascontiguousarray(thr[sh.halo_start:sh.end], np.int32)
        o_thr = np.zeros(m, np.int32); o_snr = np.zeros(m); dst = np.zeros((m, H, W), np.uint8)
        n_on = np.zeros(m, np.int32); nl = np.zeros(m, np.int32); raw = np.zeros((m, 512, 4), np.int32)
        rc = lib.emu_stream_chunk(p(part), m, C.c_longlong(sh.halo_start), halo, p(thr_in), W, H, n, case["batch"], roi, *hough,
                                  int(cfg["dy_mask"]), area, p(o_thr), p(o_snr), p(dst), p(n_on), p(nl), p(raw))
        if rc != 0:
            return f"rank {sh.rank}: chunk failed rc={rc}"
        for t in range(sh.start, sh.end):
            k = t - sh.halo_start
            if not np.array_equal(dst[k], s_dst[t]):
                return f"rank {sh.rank} frame {t}: mask differs in {int(np.count_nonzero(dst[k] != s_dst[t]))} pixels"
            if nl[k] != s_nl[t] or not np.array_equal(raw[k, :min(nl[k], 512)], s_raw[t, :min(nl[k], 512)]):
                return f"rank {sh.rank} frame {t}: segments differ"
    return None


def run_stack_case(glib, seed):
    """max_stack_kernel / gauss_stack_kernel (MaxImgContainer, FastGaussianContainer: MetLib/stacker.py:43-59) through the
    emulator against numpy with the reference's dtypes (uint16 sums and uint32 sums of squares wrap); aligned and unaligned
    buffers, chunked accumulation, clips long enough to wrap."""
    r = np.random.default_rng([seed, 5])
    fb = int(r.choice([16, 48, 160, 1000, 4096, 37, 1001])) * (1 if r.random() < 0.5 else 3)
    T = int(r.choice([1, 2, 5, 17, 64, 300]))
    off = int(r.choice([0, 0, 1, 8]))  # an unaligned base takes the byte-wise branch
    buf = np.empty(T * fb + 16, np.uint8)
    fr = buf[off:off + T * fb].reshape(T, fb)
    fr[:] = r.integers(0, 256, (T, fb), dtype=np.uint8) if r.random() < 0.5 else r.integers(200, 256, (T, fb), dtype=np.uint8)
    chunk = int(r.integers(1, T + 1))
    grid = int(r.integers(1, 9))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.full(fb, 0xEE, np.uint8)
    glib.emu_max_stack(p(fr), T, C.c_size_t(fb), chunk, grid, p(out))
    if not np.array_equal(out, fr.max(axis=0)):
        return f"max stack differs (fb={fb}, T={T}, chunk={chunk}, off={off})"
    sm = np.full(fb, 0xEEEE, np.uint16); sq = np.full(fb, 0xEEEEEEEE, np.uint32)
    glib.emu_gauss_stack(p(fr), T, C.c_size_t(fb), chunk, grid, 0, p(sm), p(sq))
    want_s = fr.astype(np.uint16).sum(axis=0, dtype=np.uint16)
    want_q = (fr.astype(np.uint32) ** 2).sum(axis=0, dtype=np.uint32)
    if not (np.array_equal(sm, want_s) and np.array_equal(sq, want_q)):
        return f"gauss stack differs (fb={fb}, T={T}, chunk={chunk}, off={off})"
    return None


def run_readback_case(glib, seed):
    """stack_readback_kernel / stack_std_kernel / window_frame_kernel (mdb_get_stack, mdb_get_std, mdb_get_window) against the
    checker's SlidingWindow (MetLib/utils.py:225-321: max, mean, sum, std with calc_std / force_int) at a random timer, windows up
    to 300 frames (more than the kernels' 256-entry pointer table), mask on the device or none."""
    from oracle import m3_oracle as O
    r = np.random.default_rng([seed, 9])
    W, H = int(r.integers(5, 40)), int(r.integers(3, 20))
    n = int(r.choice([1, 2, 5, 30, 64, 257, 300]))
    T = int(r.integers(1, min(2 * n + 3, 320)))
    R = n + int(r.integers(0, 5))
    fr = r.integers(0, 256, (T, H, W), dtype=np.uint8)
    mask = (r.random((H, W)) > 0.3).astype(np.uint8) if r.random() < 0.5 else None
    sw = O.SlidingWindow(n, (H, W))
    for f in fr:
        sw.update(f * mask if mask is not None else f)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    mx = np.zeros((H, W), np.uint8); mean = np.zeros((H, W), np.uint8); sm = np.zeros((H, W), np.uint32); newest = np.zeros((H, W), np.uint8)
    tot = C.c_ulonglong(0)
    rc = glib.emu_stack_readback(p(fr), C.c_longlong(T - 1), W, H, n, R, p(mask), int(r.integers(1, 5)), p(mx), p(mean), p(sm), C.byref(tot), p(newest))
    if rc != 0:
        return f"rc={rc}"
    tag = f"(W={W}, H={H}, n={n}, T={T}, R={R}, mask={'yes' if mask is not None else 'no'})"
    if not np.array_equal(mx, sw.max):
        return "max differs " + tag
    if not np.array_equal(mean, sw.mean):
        return "mean differs " + tag
    if not np.array_equal(sm, sw.sum):
        return "sum differs " + tag
    L = min(n, T)
    std = float(np.sqrt(tot.value / (H * W)))  # the two double operations on the host side of mdb_get_std
    if not np.isclose(std, float(sw.std), rtol=1e-12, atol=0):
        return f"std {std!r} != {float(sw.std)!r} " + tag
    if mask is not None and not np.array_equal(newest, fr[-1] * mask):
        return "window frame differs " + tag
    return None


def build_mfnr_lib(tmp):
    from emu_build import build
    lib = C.CDLL(build(tmp, "mfnr_path_emu.cpp", patched=["mfnr.cuh"], shared=True))
    lib.emu_mfnr_mix.restype = C.c_int
    return lib


def run_mfnr_case(mlib, seed):
    """mfnr_mix_stacker (MetLib/stacker.py:296-403, connect off) on a random small clip -- every mfnr.cuh kernel in the library's
    order -- against oracle/mfnr_oracle.py (pinned on golden images of the live function): at most one grey level on at most
    1e-4 of the elements (the two global means are reduced in another order), est_bg_var to 1e-12."""
    from metdetpy_b200.stacker import get_gumbel_mean
    from oracle import mfnr_oracle as MO
    r = np.random.default_rng([seed, 13])
    H, W, Ch = int(r.integers(33, 70)), int(r.integers(33, 90)), 3  # the reference function is written for colour frames
    N = int(r.choice([2, 3, 9, 16, 17, 30, 50]))
    algo = int(r.integers(0, 4))
    base = r.integers(15, 80, (H, W, Ch))
    clip = np.clip(base[None] + r.normal(0, r.uniform(1.5, 6.0), (N, H, W, Ch)), 0, 255).astype(np.uint8)
    for _ in range(int(r.integers(0, 3))):  # bright trails in single frames
        y, x0, x1 = int(r.integers(0, H - 2)), int(r.integers(0, W // 2)), int(r.integers(W // 2, W))
        clip[int(r.integers(0, N)), y:y + 2, x0:x1] = int(r.integers(180, 256))
    clip = np.ascontiguousarray(clip if Ch == 3 else clip[..., 0])
    shape = clip.shape[1:]
    out = np.zeros(shape, np.uint8)
    st = (C.c_double * 4)()
    hp, fix = float(r.choice([0.8, 0.9, 0.95])), float(r.choice([1.0, 1.5, 2.0]))
    with np.errstate(all="ignore"):
        gm = float(get_gumbel_mean(N))
    rc = mlib.emu_mfnr_mix(clip.ctypes.data_as(C.c_void_p), N, H, W, Ch, int(r.integers(1, N + 1)), algo, C.c_double(hp), 31, C.c_double(3.0),
                           C.c_double(3.0), C.c_double(3.0), C.c_double(fix), C.c_double(gm), int(N ** 0.5), out.ctypes.data_as(C.c_void_p), st)
    if rc != 0:
        return f"rc={rc}"
    name = ["mean", "sigma-clipping", "median", "med-of-med"][algo]
    ref, ost = MO.mfnr_mix(clip, highlight_preserve=hp, bg_algorithm=name, bg_fix_factor=fix, return_stats=True, backend="numpy")
    d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
    tag = f"({H}x{W}x{Ch}, N={N}, {name}, hp={hp}, fix={fix})"
    if not (d.max() <= 1 and np.count_nonzero(d) <= max(1, int(1e-4 * d.size))):
        return f"image differs: max {int(d.max())}, {int(np.count_nonzero(d))} elements " + tag
    if not np.isclose(st[0], ost["est_bg_var"], rtol=1e-12, atol=0):
        return f"est_bg_var {st[0]!r} != {ost['est_bg_var']!r} " + tag
    return None


def build_classic_lib(tmp):
    from emu_build import build_classic
    lib = C.CDLL(build_classic(tmp))
    lib.emu_classic_path.restype = C.c_int
    lib.emu_preproc.restype = C.c_int
    return lib


def run_classic_case(lib, case):
    """ClassicDetector (MetLib/Detector.py:245-299) through the emulated classic.cuh kernels + PPHT with the configured gap."""
    from metdetpy_b200.detector import select_subarea
    from oracle import classic_oracle as CO
    W, H, T, cfg, mask, fr = (case[k] for k in ("W", "H", "T", "cfg", "mask", "frames"))
    roi = (C.c_int * 4)(*[int(v) for v in select_subarea(mask, cfg["area"])])
    cap = 8192
    thr = np.zeros(T, np.int32); snr = np.zeros(T); dst = np.zeros((T, H, W), np.uint8); nl = np.zeros(T, np.int32)
    raw = np.zeros((T, cap, 4), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_classic_path(p(fr), T, W, H, case["batch"], int(cfg["adaptive"]), int(cfg["init_value"]), _SENS.index(cfg["sensitivity"]),
                              int(cfg["interval"]), roi, *[int(v) for v in cfg["hough"]], C.c_double(float(mask.sum())), p(thr), p(snr),
                              p(dst), p(nl), p(raw), cap)
    if rc != 0:
        return f"emulated path failed: rc={rc}"
    ref = CO.ClassicDetectorOracle(1.0, 25.0, mask, 10, adaptive=cfg["adaptive"], init_value=cfg["init_value"], sensitivity=cfg["sensitivity"],
                                   area=cfg["area"], interval=cfg["interval"], hough=cfg["hough"], backend="cv2")
    for t in range(T):
        ref.update(fr[t]); lines, _ = ref.detect()
        if ref.bi_threshold != thr[t]:
            return f"frame {t}: threshold {thr[t]} != {ref.bi_threshold}"
        if not np.isclose(snr[t], ref.stack.snr, rtol=1e-12, atol=0):
            return f"frame {t}: snr {snr[t]!r} != {ref.stack.snr!r}"
        if t < 3:
            continue  # no lines and no mask before four frames are there (Detector.py:264-265)
        if not np.array_equal(dst[t], ref.dst):
            return f"frame {t}: mask differs in {int(np.count_nonzero(dst[t] != ref.dst))} pixels"
        want = np.asarray(lines, np.int32).reshape(-1, 4)
        if nl[t] != len(want) or (nl[t] <= cap and not np.array_equal(raw[t, :nl[t]], want)):
            return f"frame {t}: segments differ ({nl[t]} vs {len(want)})"
    return None


def run_preproc_case(lib, seed):
    """Transform chain of the loader (MetLib/imgproc.py:70-139; resize, BGR2GRAY, mask, exposure merge) through the emulated
    preproc_kernel with the library's own tap tables, against oracle/preproc_oracle.py (which is pinned on cv2)."""
    from metdetpy_b200 import _lib
    from oracle import preproc_oracle as PO
    nat = _lib.load()
    r = np.random.default_rng([seed, 77])
    W0, H0 = int(r.integers(20, 200)), int(r.integers(12, 120))
    resize = r.random() < 0.8
    W, H = (int(r.integers(8, 2 * W0)), int(r.integers(6, 2 * H0))) if resize else (W0, H0)
    ch = 3 if r.random() < 0.8 else 1
    exp = int(r.integers(1, 5))
    T = int(r.integers(1, 9))
    fr = r.integers(0, 256, (T, H0, W0, 3) if ch == 3 else (T, H0, W0), dtype=np.uint8)
    mask = (r.random((H, W)) > 0.2).astype(np.uint8)
    taps = []
    for dst_n, src_n, clamp, stride in ((W, W0, 1, ch), (H, H0, 0, 1)):
        a = [np.zeros(dst_n, np.int32) for _ in range(4)]
        if nat.mdb_preproc_axis_taps(dst_n, src_n, clamp, *[x.ctypes.data for x in a]) != 0:
            return "mdb_preproc_axis_taps failed"
        a[0] *= stride; a[1] *= stride
        taps.append(np.ascontiguousarray(np.stack(a, 1)))
    G = (T + exp - 1) // exp
    out = np.zeros((G, H, W), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_preproc(p(fr), T, W0, H0, ch, 0, W, H, int((W0, H0) != (W, H)), exp, p(taps[0]), p(taps[1]), p(mask), p(out))
    if rc != 0:
        return f"emu_preproc rc={rc}"
    want = PO.preprocess_stream(fr, (W, H), ch == 3, mask, exp)
    if not np.array_equal(out, want):
        return f"{W0}x{H0}x{ch} -> {W}x{H}, exp {exp}, T {T}: {int(np.count_nonzero(out != want))} pixels differ"
    return None


def main():
    import tempfile
    first, count = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 50)
    lib, glib, clib = build_lib(tempfile.mkdtemp()), build_generic_lib(tempfile.mkdtemp()), build_classic_lib(tempfile.mkdtemp())
    mlib = build_mfnr_lib(tempfile.mkdtemp())
    bad = 0
    for seed in range(first, first + count):
        case = make_case(seed)
        res = [run_case(lib, case, pf) for pf in (False, True)]
        gcase = make_case(seed, any_width=True)
        res.append(run_case(glib, gcase, generic=True))
        res.append(run_classic_case(clib, gcase))
        res.append(run_preproc_case(clib, seed))
        res.append(run_stack_case(glib, seed))
        res.append(run_readback_case(glib, seed))
        res.append(run_mfnr_case(mlib, seed))
        res.append(run_sharded_case(lib, case, 2 + seed % 4))
        if seed % 3 == 0:  # the second-generation temporal kernel on every third case
            lib.emu_set_temporal_version(2)
            res.append(run_case(lib, case))
            lib.emu_set_temporal_version(3)
        if seed % 10 == 9:  # wide frames on every tenth case
            bcase, bg = make_case(seed, big=True), make_case(seed, any_width=True, big=True)
            res += [run_case(lib, bcase, pf) for pf in (False, True)] + [run_case(glib, bg, generic=True), run_classic_case(clib, bg)]
        if seed % 4 == 3:  # a dense variant of every fourth case
            dcase = make_case(seed, dense=True)
            res += [run_case(lib, dcase, pf) for pf in (False, True)]
        tag = {k: case[k] for k in ("W", "H", "n", "T", "batch")}
        print(seed, tag, case["cfg"], "generic", {k: gcase[k] for k in ("W", "H", "n", "T", "apply_mask")}, res, flush=True)
        bad += any(x not in (None, "skipped") for x in res)
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
