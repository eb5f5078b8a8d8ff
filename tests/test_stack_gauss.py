"""Streaming Gaussian stack (first half): FastGaussianContainer
(MetLib/stacker.py:52-59) = per-element sum (uint16) and sum of squares (uint32) with numpy's wrap-around.
CPU: restatement vs golden vectors of the live class.  GPU: kernel through the C ABI vs both."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import stack_oracle as SO


def _g():
    return np.load(os.path.join(GOLDEN, "gauss_stack.npz"))


def test_oracle_reproduces_reference_golden():
    g = _g()
    for name in g["names"]:
        s, q, n = SO.gauss_stack(g[f"{name}_frames"])
        assert s.dtype == np.uint16 and q.dtype == np.uint32 and n.dtype == np.int16
        assert np.array_equal(s, g[f"{name}_sum"]) and np.array_equal(q, g[f"{name}_sq"]) and np.array_equal(n, g[f"{name}_n"])
        with np.errstate(all="ignore"):
            mu, var = SO.mu_var(s, q, n)
        assert np.array_equal(mu, g[f"{name}_mu"]) and np.array_equal(var, g[f"{name}_var"], equal_nan=True)
    assert int(g["wrap300_frames"].astype(np.int64).sum(0).max()) > 65535  # the overflow case really overflows


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [7, 32, 1000])
def test_gpu_container_matches_reference_golden(chunk):
    from metdetpy_b200.stacker import FastGaussianContainer
    g = _g()
    for name in g["names"]:
        box = FastGaussianContainer(chunk=chunk)
        for f in g[f"{name}_frames"]:
            box.append(f)
        r = box.container
        assert np.array_equal(r.sum_mu, g[f"{name}_sum"]) and np.array_equal(r.square_sum, g[f"{name}_sq"])
        assert np.array_equal(r.n, g[f"{name}_n"]) and r.n.dtype == np.int16
        with np.errstate(all="ignore"):
            assert np.array_equal(r.mu, g[f"{name}_mu"]) and np.array_equal(r.var, g[f"{name}_var"], equal_nan=True)
    assert FastGaussianContainer().container is None


@pytest.mark.gpu
def test_gpu_full_size_colour_clip_against_oracle():
    """1080p colour frames (the reference stacks full-resolution BGR clips), odd byte count tail included."""
    from metdetpy_b200.stacker import FastGaussianContainer
    rng = np.random.default_rng(2)
    for shape in [(1080, 1920, 3), (37, 41, 3)]:
        frames = rng.integers(0, 256, (24,) + shape, dtype=np.uint8)
        box = FastGaussianContainer(chunk=10)
        for f in frames:
            box.append(f)
        r = box.container
        s, q, n = SO.gauss_stack(frames)
        assert np.array_equal(r.sum_mu, s) and np.array_equal(r.square_sum, q) and np.array_equal(r.n, n)
