"""Observation / action spaces of the reference's factories (observations.py:388-409, :507, :575-586, :712-741;
actions.py:125) -- gymnasium spaces when gymnasium is installed, small stand-ins with the same shapes / dtypes /
bounds otherwise (neither gymnasium nor stable_baselines3 exists in the build image)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import abi

try:  # pragma: no cover - not installed in the build container
    from gymnasium import spaces as _spaces
except Exception:  # noqa: BLE001
    _spaces = None


class Box:
    """Stand-in for gymnasium.spaces.Box (shape / dtype / bounds only)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class DictSpace:
    """Stand-in for gymnasium.spaces.Dict."""

    def __init__(self, spaces):
        self.spaces = OrderedDict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def __repr__(self):
        return f"Dict({dict(self.spaces)})"


class DiscreteSpace:
    def __init__(self, n):
        self.n = int(n)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def sample(self):
        return int(np.random.randint(0, self.n))


def observation_space(obs_kind: int, J: int, M: int, N: int):
    """The declared spaces of the reference's observation factories (observations.py:388-409, :507, :575-586,
    :712-741), as gymnasium spaces when gymnasium is installed."""
    box = _spaces.Box if _spaces else Box
    dct = _spaces.Dict if _spaces else DictSpace
    if obs_kind == abi.OBS_BINARY_ACTION:
        sp = OrderedDict(
            job_running=box(low=0, high=1, shape=(J,), dtype=np.int8),
            job_executed_on_machine=box(low=0, high=1, shape=(J, M), dtype=np.int8),
            job_progression=box(low=0, high=J, shape=(J,), dtype=np.int32),
            machine_running=box(low=0, high=1, shape=(M,), dtype=np.int8),
            machine_progression=box(low=0, high=J, shape=(M,), dtype=np.int32),
            available_jobs=box(low=0, high=1, shape=(J,), dtype=np.int8),
            current_time=box(low=0, high=1.0, shape=(1,), dtype=np.float32),
            current_transition=box(low=0, high=1, shape=(3,), dtype=np.float32))
    elif obs_kind == abi.OBS_OPERATION_ARRAY:
        sp = OrderedDict(
            operation_state=box(low=0, high=1, shape=(1, N), dtype=np.float32),
            job_locations=box(low=0, high=1, shape=(1, J), dtype=np.float32),
            current_transition=box(low=0, high=1, shape=(3,), dtype=np.float32))
    elif obs_kind == abi.OBS_TASSEL:
        keys = ("allocatable", "left_over_time", "percent_finished", "total_completion",
                "time_until_next_machine_is_free", "idle_since_last_op", "cum_idle_time")
        sp = OrderedDict((k, box(low=0, high=1, shape=(J,), dtype=np.float32)) for k in keys)
    else:
        sp = OrderedDict()
    return dct(sp)


def action_space():
    """BinaryJobActionFactory: Discrete(2, start=0) (actions.py:125)."""
    return _spaces.Discrete(2) if _spaces else DiscreteSpace(2)
