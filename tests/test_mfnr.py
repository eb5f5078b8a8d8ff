"""MFNR mix stacker (MetLib/stacker.py:296-403, connect_lines off, all four background algorithms): the CPU oracle against
golden images from the live reference (exact), the CUDA path through the C ABI against both (float64 pipeline whose two
global means are reduced in another order than numpy's: at most one grey level on at most 1e-4 of the elements, the
scalars to 1e-12)."""
import os
import types

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mfnr_oracle as MO

ALGOS = ("mean", "sigma-clipping", "median", "med-of-med")


def _cfg(algo, connect=False):
    ns = types.SimpleNamespace
    return ns(switch=True, highlight_preserve=0.9, algorithm="mfnr-mix", blur_ksize=31,
              connect_lines=ns(switch=connect, ksize_multiplier=1.5, gamma=1.0, threshold=30),
              mfnr_param=ns(bg_algorithm=algo, sigma_high=3.0, sigma_low=3.0, bg_fix_factor=1.5))


class Loader:  # the protocol _batch_stacker drives (stacker.py:146-175)
    def __init__(self, frames):
        self.frames, self.i, self.iterations, self.stopped = frames, 0, len(frames), False

    def reset(self, start_frame=None, end_frame=None):
        self.i = 0

    def start(self):
        self.i = 0

    def pop(self):
        f = self.frames[self.i]; self.i += 1
        return f

    def stop(self):
        self.stopped = True


@pytest.mark.parametrize("backend", ["cv2", "numpy"])
def test_oracle_reproduces_reference_golden(backend):
    g = np.load(os.path.join(GOLDEN, "mfnr.npz"))
    for name in g["names"]:
        for algo in ALGOS:
            mix = MO.mfnr_mix(g[f"{name}_frames"], bg_algorithm=algo, backend=backend)
            assert np.array_equal(mix, g[f"{name}_{algo}"]), (name, algo)


def test_gaussian_restatement_equals_cv2():
    import cv2
    rng = np.random.default_rng(3)
    for shape in [(40, 57), (33, 31), (90, 64)]:
        fg = (rng.random(shape) > 0.93).astype(float)
        assert np.abs(cv2.GaussianBlur(fg, (31, 31), sigmaX=3) - MO.gaussian_blur_f64(fg, 31, 3.0)).max() < 1e-15
    assert np.abs(cv2.getGaussianKernel(31, 3, cv2.CV_64F).ravel() - MO.gaussian_kernel_f64(31, 3.0)).max() < 1e-16


def _close(mix, ref):
    d = np.abs(mix.astype(np.int16) - ref.astype(np.int16))
    assert d.max() <= 1 and np.count_nonzero(d) <= max(1, int(1e-4 * d.size)), (int(d.max()), int(np.count_nonzero(d)), d.size)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ALGOS)
def test_gpu_matches_reference_golden(algo):
    from metdetpy_b200 import stacker
    g = np.load(os.path.join(GOLDEN, "mfnr.npz"))
    for name in g["names"]:
        frames = g[f"{name}_frames"]
        ld = Loader(list(frames))
        mix = stacker.mfnr_mix_stacker(ld, _cfg(algo))
        assert ld.stopped and mix.dtype == np.uint8 and mix.shape == frames.shape[1:]
        _close(mix, g[f"{name}_{algo}"])
        _, st = MO.mfnr_mix(frames, bg_algorithm=algo, return_stats=True)
        box = stacker.MfnrMixContainer(keep_frames=True, chunk=7)
        for f in frames:
            box.append(f)
        _close(box.export(0.9, 31, algo, 1.5), g[f"{name}_{algo}"])
        assert box.stats["est_bg_var"] == pytest.approx(st["est_bg_var"], rel=1e-12)
        assert box.stats["highlight_avg_diff"] == pytest.approx(st["highlight_avg_diff"], rel=1e-12)
        assert box.stats["gumbel"] == st["gumbel"]
        box.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,T", [((270, 480, 3), 40), ((61, 97, 3), 9), ((1080, 1920, 3), 12)])
def test_gpu_seeded_clips_against_oracle(shape, T):
    from metdetpy_b200 import stacker
    rng = np.random.default_rng(T)
    base = rng.integers(10, 80, shape)
    frames = np.clip(base[None] + rng.normal(0, 5.0, (T,) + shape), 0, 255).astype(np.uint8)
    H, W = shape[:2]
    for t in range(T):
        x = (11 * t) % (W - 20)
        frames[t, H // 3 + t % 5, x:x + 16] = 245
        frames[t, 3:6, 3:9] = 255
    for algo in ALGOS:
        box = stacker.MfnrMixContainer(keep_frames=algo != "mean", chunk=5)
        for f in frames:
            box.append(f)
        mix = box.export(0.9, 31, algo, 1.5)
        box.close()
        _close(mix, MO.mfnr_mix(frames, bg_algorithm=algo, backend="cv2"))


@pytest.mark.gpu
def test_gpu_refusals_and_errors():
    from metdetpy_b200 import stacker
    fr = [np.zeros((8, 8, 3), np.uint8)] * 3
    with pytest.raises(NotImplementedError):
        stacker.mfnr_mix_stacker(Loader(fr), _cfg("mean", connect=True))
    with pytest.raises(AssertionError):
        stacker.mfnr_mix_stacker(Loader(fr), _cfg("bogus"))
    assert stacker.mfnr_mix_stacker(Loader([]), _cfg("mean")) is None
    box = stacker.MfnrMixContainer(keep_frames=False)
    box.append(fr[0]); box.append(fr[1])
    with pytest.raises(Exception):
        box.export(0.9, 31, "sigma-clipping", 1.5)  # frames were not kept
    with pytest.raises(ValueError):
        box.append(np.zeros((8, 9, 3), np.uint8))
    box.close()
