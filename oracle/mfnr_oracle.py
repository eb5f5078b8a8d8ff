"""CPU restatement of the reference's MFNR mix stacker -- TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline may import it).

MetLib/stacker.py:296-403 (`mfnr_mix_stacker`) with `connect_lines.switch = False`, background algorithms "mean"
(:339-342), "sigma-clipping" (:333-338 -> `single_sigma_clipping`, :94-115), "median" and "med-of-med" (:343-349 ->
`median_of_medians`, :62-78), on top of the containers of :34-59
(`MaxImgContainer`, `AllImgContainer`, `FastGaussianContainer`) and `FastGaussianParam.mu/.var/__sub__/mask`
(MetLib/utils.py:418-509: uint16 sums and uint32 sums of squares that wrap like numpy's fixed-width arithmetic) and
`get_gumbel_mean` (:118-126).  The SNR estimates of :322-323, :399-400 only feed debug log lines and are left out.
Pinned on golden vectors from the live function (tests/golden/mfnr.npz).
backend "cv2": cv2.GaussianBlur where the reference calls it (:368-370); "numpy": a tap-by-tap restatement
(float64, BORDER_REFLECT_101), equal to cv2's to a few ulp.
"""
from __future__ import annotations

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

EULER_CONSTANT = 0.5772  # MetLib/utils.py:24


def get_gumbel_mean(n: int) -> float:  # stacker.py:118-126
    sqrt2logn = np.sqrt(2 * np.log(n))
    return (sqrt2logn - (np.log(np.log(n)) + np.log(4 * np.pi)) / (2 * sqrt2logn) + EULER_CONSTANT / sqrt2logn)


def gaussian_kernel_f64(ksize: int, sigma: float) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma, CV_64F) for ksize > 7 or sigma > 0 (no fixed small table): exp(-x^2 / 2 sigma^2),
    normalised by the sum."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return k / k.sum()


def gaussian_blur_f64(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(img(float64), (ksize, ksize), sigmaX=sigma): separable, rows first, BORDER_REFLECT_101."""
    k = gaussian_kernel_f64(ksize, sigma)
    r = ksize // 2

    def pass_1d(a, axis):
        a = np.moveaxis(a, axis, 0)
        n = a.shape[0]
        idx = np.arange(-r, n + r)
        idx = np.abs(idx)
        idx = np.where(idx >= n, 2 * (n - 1) - idx, idx)  # reflect 101 (n > r assumed)
        p = a[idx]
        out = np.zeros_like(a)
        for j in range(ksize):
            out += k[j] * p[j:j + n]
        return np.moveaxis(out, 0, axis)

    return pass_1d(pass_1d(img.astype(np.float64), 1), 0)


class Stats:
    """(sum_mu, square_sum, n) with FastGaussianParam's dtypes and formulas (utils.py:418-509)."""

    def __init__(self, sum_mu, square_sum, n, ddof=1):
        self.sum_mu, self.square_sum, self.n, self.ddof = sum_mu, square_sum, n, ddof

    @property
    def mu(self):
        return np.round(self.sum_mu / self.n)

    @property
    def var(self):
        sum_mu = np.array(self.sum_mu, dtype=self.square_sum.dtype)
        return (self.square_sum - np.square(sum_mu) / self.n) / (self.n - self.ddof)


def stack_stats(frames: np.ndarray) -> Stats:
    """FastGaussianContainer over the frames (stacker.py:52-59): uint16 / uint32 wrap-around sums, int16 count."""
    s = np.zeros(frames.shape[1:], np.uint16)
    q = np.zeros(frames.shape[1:], np.uint32)
    for f in frames:
        f16 = f.astype(np.uint16)
        s += f16
        q += np.square(f16, dtype=np.uint32)
    return Stats(s, q, np.full(frames.shape[1:], len(frames), np.int16))


def single_sigma_clipping(frames: np.ndarray, ref: Stats, sigma_high=3.0, sigma_low=3.0) -> Stats:  # stacker.py:94-115
    mu, std = ref.mu, np.sqrt(ref.var)
    hi = np.round(mu + sigma_high * std).clip(0, 255).astype(np.uint8)
    lo = np.round(mu - sigma_low * std).clip(0, 255).astype(np.uint8)
    cs = np.zeros(frames.shape[1:], np.uint16)
    cq = np.zeros(frames.shape[1:], np.uint32)
    cn = np.zeros(frames.shape[1:], np.uint16)
    for img in frames:
        m = (img > hi) | (img < lo)
        i16 = img.astype(np.uint16)
        cs += i16 * m
        cq += np.square(i16, dtype=np.uint32) * m
        cn += m.astype(np.uint16)
    return Stats(ref.sum_mu - cs, ref.square_sum - cq, ref.n - cn)


def median_of_medians(frames: np.ndarray, block_size=None) -> np.ndarray:  # stacker.py:62-78
    if block_size is None:
        block_size = int(len(frames) ** (1 / 2))
    block_num = (len(frames) - 1) // block_size + 1
    medians = [np.median(frames[i * block_size:(i + 1) * block_size], axis=0) for i in range(block_num)]
    return np.median(medians, axis=0)


def mfnr_mix(frames: np.ndarray, *, highlight_preserve=0.9, blur_ksize=31, bg_algorithm="mean", sigma_high=3.0,
             sigma_low=3.0, bg_fix_factor=1.5, backend="cv2", return_stats=False):
    """stacker.py:296-403 for (T, H, W, 3) uint8 frames, connect_lines off.  Note: :333-338 calls single_sigma_clipping
    with sigma 3.0 / 3.0 whatever the configuration says; the arguments here default to the same."""
    frames = np.asarray(frames)
    max_img = frames.max(0)
    init = stack_stats(frames)
    with np.errstate(all="ignore"):
        if bg_algorithm == "sigma-clipping":
            sc = single_sigma_clipping(frames, init, sigma_high, sigma_low)
            est_bg_mu = sc.mu
            est_bg_var = np.mean(np.sqrt(sc.var))
        elif bg_algorithm == "mean":
            est_bg_mu = init.mu
            est_bg_var = np.mean(np.sqrt(init.var))
        elif bg_algorithm in ("median", "med-of-med"):  # stacker.py:343-349
            if bg_algorithm == "median" or len(frames) <= 16:
                est_bg_mu = np.median(frames, axis=0)
            else:
                est_bg_mu = median_of_medians(frames)
            est_bg_var = np.mean(np.sqrt(init.var))
        else:
            raise NotImplementedError(bg_algorithm)
        g = get_gumbel_mean(len(frames))
        expect_max_upper = est_bg_mu + est_bg_var * g * bg_fix_factor
        diff = max_img.astype(np.float64) - expect_max_upper
        avg = np.average(diff[diff > 0])
        fg = (diff > avg) | (max_img > 255 * highlight_preserve)
        fg = np.repeat((np.sum(fg.astype(np.uint8), axis=-1) >= 1)[..., None], 3, axis=-1).astype(float)
        if backend == "cv2":
            blur = cv2.GaussianBlur(fg, ksize=(blur_ksize, blur_ksize), sigmaX=3)
        else:
            blur = np.repeat(gaussian_blur_f64(fg[..., 0], blur_ksize, 3.0)[..., None], 3, axis=-1)
        hff = 1 - ((max_img / 255 - highlight_preserve).clip(0, 1) / (1 - highlight_preserve))
        fixed = max_img.astype(np.float64) - ((est_bg_var * g) * hff)
        fixed = np.clip(fixed, 0, 255)
        mix = np.round(fixed * blur + est_bg_mu * (1 - blur)).astype(np.uint8)
    if return_stats:
        return mix, dict(est_bg_var=float(est_bg_var), gumbel=float(g), highlight_avg_diff=float(avg),
                         fg_pixels=int(fg[..., 0].sum()))
    return mix
