"""Import stub for `onnxruntime` (absent here; not on the detector hot path)."""


def set_default_logger_severity(level):
    pass


def get_available_providers():
    return ["CPUExecutionProvider"]


def get_device():
    return "CPU"


class SessionOptions:
    pass


class InferenceSession:
    def __init__(self, *a, **k):
        raise RuntimeError("onnxruntime stub: no inference in this environment")
