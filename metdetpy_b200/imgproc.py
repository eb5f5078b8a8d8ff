"""Host-side mirror of the reference's loader preprocessing (MetLib/imgproc.py:70-139 `Transform`,
MetLib/utils.py:195-204 `MergeFunction`) for uint8 frames, executed by libmetdet_b200.so on the GPU
(preproc_kernel, csrc/preproc.cuh) through the C ABI `mdb_preproc_*`.  Same method names and call
order as the reference, so `videoloader.py:300-308` reads the same:

    tr = Transform()
    tr.opencv_resize(runtime_size)      # cv2.resize(..., INTER_LINEAR)
    tr.opencv_BGR2GRAY()                # cv2.cvtColor(..., COLOR_BGR2GRAY)
    tr.mask_with(mask)                  # img * mask
    img = tr.exec_transform(frame)      # one frame, numpy in / numpy out

plus the batched form the throughput path uses: `exec_transform_many(frames, exp_frame)` = the
transform of every frame followed by `MergeFunction.max` over groups of `exp_frame` frames, with the
result optionally left on the device for `M3Detector.detect_many(..., on_device=True)`.
Results are bit-exact with cv2.  No CPU fallback: without the CUDA library / a CUDA device it raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check

INTER_LINEAR = 1  # cv2.INTER_LINEAR


class DeviceFrames:
    """(T, H, W) uint8 frames living in a preprocessing handle's device buffer (valid until the next
    exec_transform* call on the same Transform)."""

    def __init__(self, ptr: int, shape):
        self.ptr = ptr
        self.shape = tuple(shape)

    def __len__(self):
        return self.shape[0]


class MergeFunction:
    """MergeFunction.not_merge / .max (MetLib/utils.py:195-204) on the device."""

    @classmethod
    def not_merge(cls, image_stack):
        return image_stack[0]

    @classmethod
    def max(cls, image_stack):
        from . import stacker
        return stacker.merge_max(image_stack)


class Transform(object):
    """Transform (MetLib/imgproc.py:70-139): resize, BGR2GRAY / RGB2GRAY, mask_with, exec_transform."""
    MASK_FLAG = "MASK"

    def __init__(self, device: int = 0) -> None:
        self.transform: list[tuple[str, dict[str, Any]]] = []
        self.device = device
        self._handle = None
        self._key = None

    # -- builders (reference names) --------------------------------------------------------------
    def opencv_resize(self, dsize: Sequence[int], **kwargs: Any):
        interpolation = kwargs.get("resize_interpolation", INTER_LINEAR)
        if interpolation != INTER_LINEAR:
            raise NotImplementedError("device resize implements cv2.INTER_LINEAR (the reference's default)")
        self.transform.append(("resize", dict(dsize=(int(dsize[0]), int(dsize[1])))))

    def opencv_BGR2GRAY(self):
        self.transform.append(("gray", dict(rgb=False)))

    def opencv_RGB2GRAY(self):
        self.transform.append(("gray", dict(rgb=True)))

    def mask_with(self, mask: np.ndarray):
        self.transform.append(("mask", dict(mask=np.ascontiguousarray(mask, np.uint8))))

    def opencv_debayer(self, *a, **k):
        raise NotImplementedError("debayer is not part of the device preprocessing path")

    # -- plan ------------------------------------------------------------------------------------
    def _plan(self, shape):
        """Validate the step order (resize -> gray -> mask, each optional, as videoloader.py:300-308
        builds it) and return (channels, rgb, dsize, mask)."""
        H0, W0 = shape[0], shape[1]
        C_in = 1 if len(shape) == 2 else shape[2]
        order = [n for n, _ in self.transform]
        rank = {"resize": 0, "gray": 1, "mask": 2}
        if sorted(order, key=rank.get) != order or len(set(order)) != len(order):
            raise NotImplementedError(f"unsupported transform order {order}: expected resize -> gray -> mask")
        kw = dict(self.transform)
        dsize = kw.get("resize", {}).get("dsize", (W0, H0))
        if C_in == 3 and "gray" not in kw:
            raise NotImplementedError("3-channel frames need a BGR2GRAY / RGB2GRAY step (the detector path is grayscale)")
        if C_in not in (1, 3):
            raise ValueError(f"frames must have 1 or 3 channels, got {C_in}")
        mask = kw.get("mask", {}).get("mask")
        if mask is not None and mask.shape != (dsize[1], dsize[0]):
            raise ValueError(f"mask shape {mask.shape} does not match the output size {(dsize[1], dsize[0])}")
        return C_in, bool(kw.get("gray", {}).get("rgb", False)), dsize, mask

    def _get_handle(self, shape, exp_frame: int, n_out: int):
        C_in, rgb, dsize, mask = self._plan(shape)
        key = (tuple(shape), exp_frame, tuple(n for n, _ in self.transform))
        if self._handle is not None and (self._key != key or self._cap < n_out):
            self.close()
        if self._handle is None:
            lib = _lib.load()
            if lib.mdb_device_count() < 1:
                raise _lib.MetDetError("no CUDA device visible: metdetpy_b200 has no CPU fallback")
            h = C.c_void_p()
            cap = max(n_out, 1)
            check(lib.mdb_preproc_create(shape[1], shape[0], C_in, int(rgb), dsize[0], dsize[1],
                                         None if mask is None else mask.ctypes.data, exp_frame, cap,
                                         self.device, C.byref(h)), "Transform")
            self._handle, self._key, self._cap, self._dsize = h, key, cap, dsize
        return self._handle

    # -- execution -------------------------------------------------------------------------------
    def exec_transform(self, img: np.ndarray) -> np.ndarray:
        """One frame through the chain (reference signature)."""
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8:
            raise ValueError("device preprocessing handles uint8 frames")
        return self.exec_transform_many(img[None], 1)[0]

    def exec_transform_many(self, frames, exp_frame: int = 1, *, on_device: bool = False,
                            keep_on_device: bool = False, shape=None):
        """frames: (T,H0,W0[,3]) uint8 numpy array, or (device_ptr, T) with on_device=True and
        shape=(H0,W0[,3]).  Returns the (ceil(T/exp_frame), H, W) uint8 result as a numpy array, or --
        keep_on_device=True -- as DeviceFrames for detect_many(on_device=True)."""
        lib = _lib.load()
        if on_device:
            ptr, T = frames
            fshape = tuple(shape)
        else:
            frames = np.ascontiguousarray(frames)
            if frames.dtype != np.uint8 or frames.ndim not in (3, 4):
                raise ValueError("frames must be a (T,H,W) or (T,H,W,3) uint8 array")
            T, fshape, ptr = len(frames), frames.shape[1:], frames.ctypes.data
        if T < 1:
            raise ValueError("no frames")
        exp_frame = int(exp_frame)
        G = (T + exp_frame - 1) // exp_frame
        h = self._get_handle(fshape, exp_frame, G)
        W, H = self._dsize
        n = C.c_int32()
        if keep_on_device:
            check(lib.mdb_preproc_run(h, ptr, T, int(on_device), None, 0, C.byref(n)), "exec_transform_many")
            p = C.c_void_p()
            check(lib.mdb_preproc_output(h, C.byref(p)), "exec_transform_many")
            return DeviceFrames(p.value, (n.value, H, W))
        out = np.empty((G, H, W), np.uint8)
        check(lib.mdb_preproc_run(h, ptr, T, int(on_device), out.ctypes.data, 0, C.byref(n)), "exec_transform_many")
        return out

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(_lib.load().mdb_preproc_time(self._handle, C.byref(ms)), "Transform")
        return ms.value

    def close(self):
        if self._handle is not None:
            _lib.load().mdb_preproc_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
