"""CPU restatement of the reference's streaming Gaussian stack -- TEST INFRASTRUCTURE ONLY.

FastGaussianContainer.append (MetLib/stacker.py:52-59) wraps every frame into
FastGaussianParam(frame.astype(uint16)) (MetLib/utils.py:435-452: sum_mu = the uint16 frame,
square_sum = np.square(sum_mu, dtype=uint32), n = ones int16) and adds it to the running one
(__add__, utils.py:485-493).  numpy keeps the dtypes, so after T frames
    sum_mu     = (sum_t x_t)   mod 2^16   (uint16)
    square_sum = (sum_t x_t^2) mod 2^32   (uint32)
    n          = T wrapped to int16
and mu = round(sum_mu / n), var = (square_sum - sum_mu^2 / n) / (n - ddof) (utils.py:454-465).
Pinned on golden vectors from the live class (tests/golden/gauss_stack.npz), overflow case included.
"""
from __future__ import annotations

import numpy as np


def gauss_stack(frames: np.ndarray):
    """(T, ...) uint8 -> (sum_mu uint16, square_sum uint32, n int16) of FastGaussianContainer after T appends."""
    f = np.asarray(frames, np.uint8)
    sum_mu = np.add.reduce(f.astype(np.uint16), axis=0, dtype=np.uint16)
    sq = np.add.reduce(np.square(f.astype(np.uint16), dtype=np.uint32), axis=0, dtype=np.uint32)
    n = np.full(f.shape[1:], np.array(len(f)).astype(np.int64).astype(np.int16), np.int16)
    return sum_mu, sq, n


def mu_var(sum_mu, square_sum, n, ddof: int = 1):
    mu = np.round(sum_mu / n)
    s = np.array(sum_mu, dtype=square_sum.dtype)
    var = (square_sum - np.square(s) / n) / (n - ddof)
    return mu, var
