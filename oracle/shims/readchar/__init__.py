"""Import shim (test infrastructure only): permissive dummy module."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _dummy import Dummy  # noqa: E402


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Dummy()
