"""Config dataclasses with the reference's field names (MetLib/metstruct.py:365-389).

The detector accepts either these or the reference's own `BinaryCfg` object (duck-typed), so that
`MetDetPy.detect_video` can hand over `cfg.detector.cfg` unchanged (MetDetPy.py:137-142)."""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass
class BinaryCoreCfg:
    adaptive_bi_thre: bool = True
    init_value: int = 7
    sensitivity: str = "normal"
    area: float = 0.1
    interval: int = 2


@dataclasses.dataclass
class HoughLineCfg:
    threshold: int = 10
    min_len: int = 10
    max_gap: int = 10


@dataclasses.dataclass
class DynamicCfg:
    dy_mask: bool = True
    window_sec: float = 5  # parsed but never used by the reference


@dataclasses.dataclass
class BinaryCfg:
    binary: BinaryCoreCfg = dataclasses.field(default_factory=BinaryCoreCfg)
    hough_line: HoughLineCfg = dataclasses.field(default_factory=HoughLineCfg)
    dynamic: DynamicCfg = dataclasses.field(default_factory=DynamicCfg)
