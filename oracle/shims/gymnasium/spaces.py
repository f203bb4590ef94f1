"""Shim for gymnasium.spaces: only what jobshoplab touches."""
import random as _random

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype

    def contains(self, x):
        return True

    def sample(self):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, start=0, seed=None):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)
        self._rng = _random.Random(seed)

    def contains(self, x):
        if isinstance(x, (bool, np.bool_)):
            return False
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self):
        return self.start + self._rng.randrange(self.n)

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, o):
        return isinstance(o, Discrete) and (o.n, o.start) == (self.n, self.start)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
        self.low = low
        self.high = high

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"

    def __eq__(self, o):
        return (
            isinstance(o, Box)
            and self.shape == o.shape
            and np.dtype(self.dtype) == np.dtype(o.dtype)
            and np.all(np.asarray(self.low) == np.asarray(o.low))
            and np.all(np.asarray(self.high) == np.asarray(o.high))
        )


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def __eq__(self, o):
        return isinstance(o, Dict) and self.spaces == o.spaces

    def __repr__(self):
        return f"Dict({self.spaces})"


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        super().__init__((n,), np.int8)
        self.n = n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec)
        super().__init__(self.nvec.shape, dtype)
