"""ctypes binding of libmetdet_b200.so (the C ABI in include/metdet_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc (metdetpy_b200/build.py);
if that is impossible, or no CUDA device is present when a detector is created, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

from . import build as _build

MAX_LINES = 512
NUM_LINES_TOOMUCH = 500
ABI_VERSION = 102  # mdb_version() of the library this binding was written against


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("window", C.c_int32),
                ("adaptive", C.c_int32), ("init_value", C.c_int32), ("sensitivity", C.c_int32),
                ("nz_interval", C.c_int32), ("roi", C.c_int32 * 4), ("hough_threshold", C.c_int32),
                ("hough_min_len", C.c_int32), ("hough_max_gap", C.c_int32), ("dy_mask", C.c_int32),
                ("max_batch", C.c_int32), ("device", C.c_int32), ("apply_mask", C.c_int32),
                ("detector", C.c_int32), ("reserved", C.c_int32 * 3)]


class MfnrParams(C.Structure):
    _fields_ = [("highlight_preserve", C.c_double), ("blur_ksize", C.c_int32), ("bg_algorithm", C.c_int32),
                ("blur_sigma", C.c_double), ("sigma_high", C.c_double), ("sigma_low", C.c_double),
                ("bg_fix_factor", C.c_double), ("gumbel_mean", C.c_double), ("med_block_size", C.c_int32),
                ("reserved", C.c_int32)]


class FrameInfo(C.Structure):
    _fields_ = [("timer", C.c_int64), ("bi_threshold", C.c_int32), ("n_on", C.c_int32),
                ("bi_threshold_float", C.c_double), ("snr", C.c_double), ("dst_sum", C.c_double),
                ("gap", C.c_double), ("lines_num", C.c_int32), ("n_raw", C.c_int32),
                ("n_lines", C.c_int32), ("len_ties", C.c_int32)]


# every symbol include/metdet_b200.h declares: (restype, argtypes)
_VP, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = {
    "mdb_last_error": (C.c_char_p, []),
    "mdb_version": (_I, []),
    "mdb_device_count": (_I, []),
    "mdb_create": (_I, [C.POINTER(Config), _VP, C.POINTER(_VP)]),
    "mdb_destroy": (_I, [_VP]),
    "mdb_update": (_I, [_VP, _VP, _I]),
    "mdb_detect": (_I, [_VP, C.POINTER(FrameInfo), _VP, _VP, _VP]),
    "mdb_detect_batch": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _I]),
    "mdb_submit_batch": (_I, [_VP, _VP, _I, _I]),
    "mdb_collect_batch": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I]),
    "mdb_get_dst": (_I, [_VP, _VP, _I]),
    "mdb_get_dst_device": (_I, [_VP, C.POINTER(_VP)]),
    "mdb_get_stack": (_I, [_VP, _VP, _VP, _VP]),
    "mdb_get_window": (_I, [_VP, _VP, _I]),
    "mdb_get_std": (_I, [_VP, C.POINTER(C.c_double)]),
    "mdb_get_raw_lines": (_I, [_VP, _I, _VP, _I, C.POINTER(C.c_int32)]),
    "mdb_get_stream": (_I, [_VP, C.POINTER(_VP)]),
    "mdb_get_launch_count": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mdb_get_fused_time": (_I, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "mdb_max_stack": (_I, [_VP, _I, _SZ, _VP, _I, _I, _I]),
    "mdb_lineset_nms": (_I, [_VP, _I, _VP, _VP, C.POINTER(C.c_int32)]),
    "mdb_lineset_nms_ordered": (_I, [_VP, _I, _VP, _VP, _VP, C.POINTER(C.c_int32)]),
    "mdb_lineset_nms_frames": (_I, [_I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mdb_alloc_pinned": (_I, [_SZ, C.POINTER(_VP)]),
    "mdb_free_pinned": (_I, [_VP]),
    "mdb_set_option": (_I, [_VP, C.c_char_p, _I]),
    "mdb_get_info": (_I, [_VP, C.c_char_p, C.POINTER(C.c_double)]),
    "mdb_reset": (_I, [_VP]),
    "mdb_noise_sums_dev": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "mdb_replay_thresholds": (_I, [_I, _VP, _VP, C.c_int64, _I, _I, _I, _I, _I, C.c_int64, C.c_int64, _VP, _VP, _VP]),
    "mdb_debug_timeline": (_I, [_VP, _VP]),
    "mdb_debug_hough_profile": (_I, [_VP, _VP, _I]),
    "mdb_seek": (_I, [_VP, C.c_int64]),
    "mdb_noise_sums": (_I, [_VP, _I, _I, C.c_int64, _I, _I, _I, _I, _VP, _VP, _VP, _I]),
    "mdb_submit_batch_thr": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP]),
    "mdb_submit_batch_ex": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _I]),
    "mdb_gauss_stack": (_I, [_VP, _I, _SZ, _VP, _VP, _I, _I, _I, _I]),
    "mdb_mfnr_create": (_I, [_I, _I, _I, _I, _I, C.POINTER(_VP)]),
    "mdb_mfnr_reserve": (_I, [_VP, _I]),
    "mdb_mfnr_append": (_I, [_VP, _VP, _I, _I]),
    "mdb_mfnr_finish": (_I, [_VP, C.POINTER(MfnrParams), _VP, _I, _VP]),
    "mdb_mfnr_stats": (_I, [_VP, _VP, _VP, _VP, C.POINTER(C.c_int64)]),
    "mdb_mfnr_destroy": (_I, [_VP]),
    "mdb_preproc_create": (_I, [_I, _I, _I, _I, _I, _I, _VP, _I, _I, _I, C.POINTER(_VP)]),
    "mdb_preproc_run": (_I, [_VP, _VP, _I, _I, _VP, _I, C.POINTER(C.c_int32)]),
    "mdb_preproc_output": (_I, [_VP, C.POINTER(_VP)]),
    "mdb_preproc_time": (_I, [_VP, C.POINTER(C.c_float)]),
    "mdb_preproc_destroy": (_I, [_VP]),
    "mdb_preproc_axis_taps": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the CUDA library. Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:
            if not os.path.exists(path):
                raise RuntimeError(
                    "libmetdet_b200.so is missing and could not be built with nvcc "
                    f"({e}); metdetpy_b200 has no CPU fallback") from e
            # an older binary exists: it is only usable if it still speaks this binding's ABI (checked below)
            warnings.warn(f"libmetdet_b200.so is older than its sources and the rebuild failed ({e}); "
                          "loading the existing binary", RuntimeWarning, stacklevel=2)
    lib = C.CDLL(path)
    lib.mdb_version.restype = _I
    ver = lib.mdb_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"{path} reports ABI version {ver}, this binding needs {ABI_VERSION}: rebuild it "
                           "(python -m metdetpy_b200.build --force)")
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise RuntimeError(f"{path} does not export {name} (declared in include/metdet_b200.h): stale build")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MetDetError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mdb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise MetDetError(f"{what}: {msg} (code {rc})")
