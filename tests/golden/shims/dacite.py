"""Minimal stand-in for the `dacite` package (not installed in this image).

Test infrastructure only: lets the *unmodified* reference (`/root/reference/MetLib`)
import in the build container so golden vectors can be generated from it
(`tests/golden/make_golden.py`). Handles what the reference's config dataclasses
need: nested dataclasses, Optional/Union, list[...], dict, int->float.
"""
import dataclasses
import typing


def _build(tp, value):
    origin = typing.get_origin(tp)
    if tp is typing.Any:
        return value
    if dataclasses.is_dataclass(tp):
        if isinstance(value, tp):
            return value
        if not isinstance(value, dict):
            raise TypeError(f"expected dict for {tp}, got {type(value)}")
        return from_dict(tp, value)
    if origin is typing.Union:
        args = typing.get_args(tp)
        if value is None and type(None) in args:
            return None
        last = None
        for a in args:
            if a is type(None):
                continue
            try:
                return _build(a, value)
            except Exception as e:  # try the next member
                last = e
        raise TypeError(f"no member of {tp} accepts {value!r}: {last}")
    if origin in (list, typing.List):
        (a,) = typing.get_args(tp) or (typing.Any,)
        return [_build(a, v) for v in value]
    if origin in (tuple, typing.Tuple):
        return tuple(value)
    if origin in (dict, typing.Dict):
        return dict(value)
    if tp is float and isinstance(value, int) and not isinstance(value, bool):
        return float(value)
    if isinstance(tp, type) and not isinstance(value, tp):
        raise TypeError(f"expected {tp}, got {type(value)} ({value!r})")
    return value


def from_dict(data_class, data, config=None):
    hints = typing.get_type_hints(data_class)
    kwargs = {}
    for f in dataclasses.fields(data_class):
        if not f.init:
            continue
        if f.name in data:
            kwargs[f.name] = _build(hints[f.name], data[f.name])
        elif f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING:
            raise KeyError(f"missing field {f.name} for {data_class.__name__}")
    return data_class(**kwargs)
