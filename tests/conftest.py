import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

DET_CASES = ["synth_320x240_n5_dyoff", "synth_384x216_n12_dyon_mask", "synth_203x157_n3_high",
             "synth_256x160_n6_fixed3_dense", "synth_300x200_n7_low", "clip_192x144_n25"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_det_case(name):
    g = np.load(os.path.join(GOLDEN, f"det_{name}.npz"))
    d = {k: g[k] for k in g.files}
    T, H, W = d["frames"].shape
    d["dst"] = np.unpackbits(d["dst_bits"], axis=1)[:, :H * W].reshape(T, H, W) * np.uint8(255)
    d["cfg"] = dict(adaptive=bool(d["cfg_adaptive"]), init_value=int(d["cfg_init_value"]),
                    sensitivity=str(d["cfg_sensitivity"]), area=float(d["cfg_area"]),
                    interval=int(d["cfg_interval"]), hough=tuple(int(x) for x in d["hough"]),
                    dy_mask=bool(d["dy_mask"]))
    d["n"] = int(d["n"]); d["fps"] = float(d["fps"])
    return d


def ragged_get(flat, offs, i):
    return flat[int(offs[i]):int(offs[i + 1])]


def has_len2_ties(lines):
    if len(lines) < 2:
        return False
    l2 = (lines[:, 2] - lines[:, 0]) ** 2 + (lines[:, 3] - lines[:, 1]) ** 2
    return len(np.unique(l2)) != len(l2)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
