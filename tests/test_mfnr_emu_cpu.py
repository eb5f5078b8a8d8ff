"""The frame-pass kernels of the MFNR stacker (csrc/mfnr.cuh: accumulation, sigma clipping, medians) on the HOST:
tests/emu/mfnr_host_emu.cpp includes the kernel source with the CUDA built-ins emulated and checks every element against
brute-force statements of the reference's containers (MetLib/stacker.py:43-59, uint16 / uint32 wrap-around incl. a 300-frame
clip that wraps), single_sigma_clipping (:94-115) and np.median / median_of_medians (:62-78).  No GPU needed."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mfnr_frame_pass_kernels_against_brute_force(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", "/nonexistent"), "include"))
                     if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "mfnr_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", cuda_inc,
                           os.path.join(REPO, "tests", "emu", "mfnr_host_emu.cpp"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.fixture(scope="module")
def mfnr_path_lib(tmp_path_factory):
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from emu_build import build
    lib = C.CDLL(build(tmp_path_factory.mktemp("mfnr_path"), "mfnr_path_emu.cpp", patched=["mfnr.cuh"], shared=True))
    lib.emu_mfnr_mix.restype = C.c_int
    return lib


@pytest.mark.parametrize("algo,code", [("mean", 0), ("sigma-clipping", 1), ("median", 2), ("med-of-med", 3)])
def test_whole_mfnr_mix_with_emulated_kernels_against_reference_golden(mfnr_path_lib, algo, code):
    """Every kernel of csrc/mfnr.cuh (frame passes, the two deterministic reductions, mask, separable Gaussian, mix) run by
    the thread-block emulator in mdb_mfnr_append / mdb_mfnr_finish's order, fed chunk-wise like MfnrMixContainer, against golden
    images of the live mfnr_mix_stacker (MetLib/stacker.py:296-403) -- the bar tests/test_mfnr.py holds on the GPU."""
    import ctypes as C

    import numpy as np
    from metdetpy_b200.stacker import get_gumbel_mean
    from oracle import mfnr_oracle as MO
    lib = mfnr_path_lib
    g = np.load(os.path.join(REPO, "tests", "golden", "mfnr.npz"))
    for name in g["names"]:  # 13 frames: plain median; 50 frames: med-of-med really runs in blocks
        frames = np.ascontiguousarray(g[f"{name}_frames"])
        N, H, W, Ch = frames.shape
        out = np.zeros((H, W, Ch), np.uint8)
        st = (C.c_double * 4)()
        rc = lib.emu_mfnr_mix(frames.ctypes.data_as(C.c_void_p), N, H, W, Ch, 7, code, C.c_double(0.9), 31, C.c_double(3.0),
                              C.c_double(3.0), C.c_double(3.0), C.c_double(1.5), C.c_double(float(get_gumbel_mean(N))),
                              int(N ** 0.5), out.ctypes.data_as(C.c_void_p), st)
        assert rc == 0, rc
        ref = g[f"{name}_{algo}"]
        d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1 and np.count_nonzero(d) <= max(1, int(1e-4 * d.size)), (name, int(d.max()), int(np.count_nonzero(d)))
        _, ost = MO.mfnr_mix(frames, bg_algorithm=algo, return_stats=True)
        assert abs(st[0] - ost["est_bg_var"]) <= 1e-12 * abs(ost["est_bg_var"]), (st[0], ost["est_bg_var"])
