import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

DET_CASES = ["synth_320x240_n5_dyoff", "synth_384x216_n12_dyon_mask", "synth_203x157_n3_high",
             "synth_256x160_n6_fixed3_dense", "synth_300x200_n7_low", "clip_192x144_n25",
             # BASELINE config 1: the bundled clip as detect_video feeds it (real mask, exp_frame = 4 merge, n = 6)
             "clip_cfg1_480x270_n6", "clip_cfg1_960x540_n6_range"]


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


def assert_nms_equivalent(got_lines, got_prob, ref_lines, ref_prob, raw_lines, ctx=""):
    """NMS output check. Without length ties among the raw segments the result must be identical
    (rows, order, probabilities to 1e-12). With ties the reference's own order comes from an
    unstable np.argsort (MetLib/utils.py:804) and depends on numpy's CPU dispatch, so the kept
    segments are compared as a set and, failing that (a tie decided which of two mutually
    absorbing segments survives), by count."""
    got = np.asarray(got_lines).reshape(-1, 4)
    ref = np.asarray(ref_lines).reshape(-1, 4)
    raw = np.asarray(raw_lines).reshape(-1, 4)
    if not has_len2_ties(raw):
        assert np.array_equal(got, ref), (ctx, got, ref)
        if got_prob is not None and ref_prob is not None:
            assert np.allclose(np.asarray(got_prob).ravel(), np.asarray(ref_prob).ravel(), rtol=1e-12, atol=0), ctx
        return
    a = sorted(map(tuple, got.tolist()))
    b = sorted(map(tuple, ref.tolist()))
    if a != b:
        assert abs(len(a) - len(b)) <= 2, (ctx, got, ref)
