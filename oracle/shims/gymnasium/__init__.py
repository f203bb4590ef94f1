"""Import shim (test infrastructure only): minimal stand-in for `gymnasium`
so the UNMODIFIED reference at /root/reference imports in this container,
where gymnasium is not installed and there is no network.  Never shipped,
never imported by the product package."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass


Space = spaces.Space
