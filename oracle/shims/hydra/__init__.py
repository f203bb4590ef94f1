"""Shim for hydra (test infrastructure only): `initialize` + `compose` as
used by jobshoplab/utils/load_config.py:15-46 -- yaml.safe_load of
<config_path>/<config_name>.yaml resolved relative to the caller's file."""
import inspect
import os
from contextlib import contextmanager

import yaml

from . import utils  # noqa: F401

_STATE = {"dir": None}


@contextmanager
def initialize(version_base=None, config_path=None, job_name=None):
    caller = inspect.stack()[2].filename
    base = os.path.dirname(os.path.abspath(caller))
    _STATE["dir"] = os.path.normpath(os.path.join(base, config_path or "."))
    try:
        yield
    finally:
        _STATE["dir"] = None


def _set(d, dotted, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = yaml.safe_load(value)


def compose(config_name=None, overrides=None):
    path = os.path.join(_STATE["dir"], config_name)
    if not path.endswith((".yaml", ".yml")):
        path += ".yaml"
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    for ov in overrides or []:
        k, v = ov.split("=", 1)
        _set(cfg, k.lstrip("+"), v)
    return cfg
