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
    """NMS output check: identical rows, order and probabilities (1e-12).  The reference orders the raw segments with
    np.argsort(len^2)[::-1] (MetLib/utils.py:804); among EQUAL lengths that order is numpy's / the CPU's business (an
    insertion sort on some hosts, a SIMD sorting network with a different tie order on AVX-512 / AVX2 hosts -- the
    GPU box is not the machine the golden files were made on).  For a frame with such a tie the expected result is
    recomputed on THIS host from the golden raw segments with the CPU checker's NMS, which calls numpy's argsort
    exactly like the reference -- and the comparison is exact again."""
    got = np.asarray(got_lines).reshape(-1, 4)
    ref = np.asarray(ref_lines).reshape(-1, 4)
    raw = np.asarray(raw_lines).reshape(-1, 4)
    if has_len2_ties(raw):
        from oracle import m3_oracle as O
        ref, ref_prob = O.lineset_nms(raw.astype(np.int32))
        NMS_BRANCHES["recomputed_on_host"] += 1
    else:
        NMS_BRANCHES["golden_exact"] += 1
    assert np.array_equal(got, np.asarray(ref).reshape(-1, 4)), (ctx, got, ref)
    if got_prob is not None and ref_prob is not None:
        assert np.allclose(np.asarray(got_prob).ravel(), np.asarray(ref_prob).ravel(), rtol=1e-12, atol=0), ctx


NMS_BRANCHES = {"golden_exact": 0, "recomputed_on_host": 0}
