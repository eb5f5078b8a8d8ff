"""Import stub for PyAV (absent here; decode is out of scope)."""
from . import error  # noqa: F401


def open(*a, **k):
    raise RuntimeError("av stub: PyAV is not available")


class VideoFrame:
    pass
