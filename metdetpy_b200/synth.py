"""Deterministic synthetic night-sky frame stream (test / bench input, not product logic).

This is the generator specified in SURVEY.md Appendix E / §8(d): static sky at grey level 48 with
blurred stars, per-frame Gaussian sensor noise (sigma = 2 grey levels) and one anti-aliased streak
("meteor") every 2 s lasting 0.5 s.  The reference's behaviour on it is documented there (threshold
settles at 8, sparse masks, one NMS line per meteor frame).  Host-side numpy/cv2 only; the frames
are uploaded to the GPU path and fed to the CPU checker unchanged so both see identical bytes.

`speed_scale` / `thickness` exist so that small test-sized streams still contain streaks long
enough for HoughLinesP(minLineLength=10); the defaults reproduce Appendix E exactly.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is part of the image; the generator needs GaussianBlur and line(LINE_AA)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def make_sky(W: int, H: int, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    sky = np.full((H, W), 24.0, np.float32)
    ns = int(W * H / 20000)
    sx = rng.integers(0, W, ns)
    sy = rng.integers(0, H, ns)
    sa = rng.uniform(30, 200, ns)  # draw order matters
    for x, y, a in zip(sx, sy, sa):
        sky[y, x] += a
    sky = cv2.GaussianBlur(sky, (0, 0), 1.2) * 4
    return np.clip(sky - 72, 0, None) + 24


def make_frame(t: int, sky: np.ndarray, W: int, H: int, FPS: float, seed: int = 1234,
               speed_scale: float = 1.0, thickness: int | None = None,
               sigma: float = 2.0) -> np.ndarray:
    f = sky + np.random.default_rng([seed, t]).normal(0, sigma, (H, W)).astype(np.float32)
    period, dur = int(2 * FPS), int(0.5 * FPS)
    k, ph = t // period, t % period
    if ph < dur:
        rr = np.random.default_rng([seed, 10**6 + k])
        x0 = rr.uniform(0.2, 0.8) * W
        y0 = rr.uniform(0.2, 0.8) * H
        ang = rr.uniform(0, 2 * np.pi)
        v = 12 * W / 1920 * 30 / FPS * speed_scale
        p1 = (int(x0 + v * ph * np.cos(ang)), int(y0 + v * ph * np.sin(ang)))
        p2 = (int(x0 + v * (ph + 1) * np.cos(ang)), int(y0 + v * (ph + 1) * np.sin(ang)))
        m = np.zeros((H, W), np.float32)
        th = max(1, int(2 * W / 1920)) if thickness is None else thickness
        cv2.line(m, p1, p2, 60.0, th, cv2.LINE_AA)
        f += m
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def make_stream(T: int, W: int, H: int, FPS: float, seed: int = 1234, t0: int = 0,
                speed_scale: float = 1.0, thickness: int | None = None,
                sigma: float = 2.0, out: np.ndarray | None = None) -> np.ndarray:
    """Frames t0 .. t0+T-1 as one (T,H,W) uint8 array."""
    sky = make_sky(W, H, seed)
    if out is None:
        out = np.empty((T, H, W), np.uint8)
    for i in range(T):
        out[i] = make_frame(t0 + i, sky, W, H, FPS, seed, speed_scale, thickness, sigma)
    return out


def make_stream_device(T: int, W: int, H: int, FPS: float, device, seed: int = 1234, t0: int = 0,
                       loop: int = 0, quiet: int = 0):
    """Same distribution as make_stream, produced on the GPU with torch (Philox noise) for
    throughput runs where host generation of 4K/8K streams would dominate (SURVEY App. E allows this;
    parity runs use the host generator). Returns a (T,H,W) uint8 CUDA tensor.

    loop > 0 makes the stream cyclic with period `loop` frames (a benchmark replays a few resident
    batches).  quiet > 0 (not used by bench.py any more) leaves out a streak that would still be inside
    the detector's window (`quiet` frames) when the stream jumps back to frame 0."""
    import torch
    sky = torch.from_numpy(make_sky(W, H, seed)).to(device)
    out = torch.empty((T, H, W), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    period, dur = int(2 * FPS), int(0.5 * FPS)
    for i in range(T):
        t = t0 + i
        g.manual_seed(seed * 1000003 + t)
        f = sky + torch.randn((H, W), generator=g, device=device, dtype=torch.float32) * 2.0
        tm = t % loop if loop > 0 else t
        k, ph = tm // period, tm % period
        if loop > 0 and k * period + dur + quiet > loop:
            ph = dur  # this streak would straddle the wrap: not drawn
        if ph < dur:
            rr = np.random.default_rng([seed, 10**6 + k])
            x0 = rr.uniform(0.2, 0.8) * W
            y0 = rr.uniform(0.2, 0.8) * H
            ang = rr.uniform(0, 2 * np.pi)
            v = 12 * W / 1920 * 30 / FPS
            p1 = (int(x0 + v * ph * np.cos(ang)), int(y0 + v * ph * np.sin(ang)))
            p2 = (int(x0 + v * (ph + 1) * np.cos(ang)), int(y0 + v * (ph + 1) * np.sin(ang)))
            th = max(1, int(2 * W / 1920))
            pad = th + 3
            bx0, by0 = max(0, min(p1[0], p2[0]) - pad), max(0, min(p1[1], p2[1]) - pad)
            bx1, by1 = min(W, max(p1[0], p2[0]) + pad + 1), min(H, max(p1[1], p2[1]) + pad + 1)
            if bx1 > bx0 and by1 > by0:
                m = np.zeros((by1 - by0, bx1 - bx0), np.float32)
                cv2.line(m, (p1[0] - bx0, p1[1] - by0), (p2[0] - bx0, p2[1] - by0), 60.0, th, cv2.LINE_AA)
                f[by0:by1, bx0:bx1] += torch.from_numpy(m).to(device)
        out[i] = torch.clamp(torch.round(f), 0, 255).to(torch.uint8)
    return out
