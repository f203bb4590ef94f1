"""Shim for omegaconf (test infrastructure only): structured/merge/to_object
building the reference's Config dataclass from a plain dict."""
import dataclasses
import typing


class DictConfig(dict):
    pass


def _build(cls, data):
    if not dataclasses.is_dataclass(cls):
        return data
    hints = typing.get_type_hints(cls)
    kwargs = {}
    for f in dataclasses.fields(cls):
        if data is None or f.name not in data:
            continue
        v = data[f.name]
        t = hints.get(f.name, None)
        if dataclasses.is_dataclass(t) and isinstance(v, dict):
            kwargs[f.name] = _build(t, v)
        elif t is float and isinstance(v, int) and not isinstance(v, bool):
            kwargs[f.name] = float(v)
        elif t is tuple or typing.get_origin(t) is tuple:
            kwargs[f.name] = tuple(v) if v is not None else v
        else:
            kwargs[f.name] = v
    return cls(**kwargs)


class _Structured:
    def __init__(self, cls):
        self.cls = cls


class OmegaConf:
    @staticmethod
    def structured(cls):
        return _Structured(cls)

    @staticmethod
    def create(d=None):
        return dict(d or {})

    @staticmethod
    def merge(*cfgs):
        cls = None
        data = {}

        def deep(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    deep(a[k], v)
                else:
                    a[k] = v

        for c in cfgs:
            if isinstance(c, _Structured):
                cls = c.cls
            elif c:
                deep(data, dict(c))
        return (cls, data)

    @staticmethod
    def to_object(merged):
        cls, data = merged
        return _build(cls, data) if cls else data

    @staticmethod
    def to_container(c, **k):
        return c
