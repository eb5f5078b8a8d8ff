"""metdetpy_b200 -- B200 (sm_100a) implementation of MetDetPy's M3 line-detector hot path.

Drop-in classes: `M3Detector`, `ClassicDetector`, `LineDetector`, `SlidingWindow`, `SNR_SW`, `EMA`,
`lineset_nms` (MetLib/Detector.py, MetLib/utils.py), `max_stacker`, `MaxImgContainer` (MetLib/stacker.py)
and the loader's `Transform` / `MergeFunction` (MetLib/imgproc.py, MetLib/utils.py; module `imgproc`).
The CUDA library is loaded on first use; there is no CPU fallback."""
from .config import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg  # noqa: F401

__all__ = ["M3Detector", "ClassicDetector", "Transform", "LineDetector", "BaseDetector", "SlidingWindow", "SNR_SW", "EMA",
           "lineset_nms", "select_subarea", "max_stacker", "MaxImgContainer", "merge_max",
           "BinaryCfg", "BinaryCoreCfg", "HoughLineCfg", "DynamicCfg", "register_with_metlib"]


def __getattr__(name):
    if name == "Transform":
        from . import imgproc
        return imgproc.Transform
    if name in ("M3Detector", "ClassicDetector", "LineDetector", "BaseDetector", "SlidingWindow", "SNR_SW", "EMA",
                "lineset_nms", "select_subarea"):
        from . import detector
        return getattr(detector, name)
    if name in ("max_stacker", "MaxImgContainer", "merge_max"):
        from . import stacker
        return getattr(stacker, name)
    raise AttributeError(name)


def register_with_metlib():
    """Plug the CUDA detector into the reference's registry (MetLib/__init__.py:34-46) so that
    `get_detector("M3Detector")` -- i.e. an unmodified MetDetPy.detect_video -- returns it."""
    import sys

    import MetLib  # the reference package must be importable
    import MetLib.Detector as D

    from .detector import ClassicDetector, M3Detector
    D.M3Detector = M3Detector
    D.ClassicDetector = ClassicDetector
    MetLib.M3Detector = M3Detector
    ours = {"M3Detector": M3Detector, "ClassicDetector": ClassicDetector}
    MetLib.available_detectors = [ours.get(getattr(c, "__name__", ""), c) for c in MetLib.available_detectors]
    # get_xxx closes over a dict built at import time (MetLib/__init__.py:16-25): rebuild it
    MetLib.get_detector = MetLib.get_xxx("detector", MetLib.available_detectors)
    main = sys.modules.get("MetDetPy")
    if main is not None and hasattr(main, "get_detector"):
        main.get_detector = MetLib.get_detector
        main.M3Detector = M3Detector
    return M3Detector
