"""Sample tables of stochastic time objects.

The reference draws every sample from a FRESH ``numpy.random.default_rng(seed)`` whose seed starts at
``np.random.randint(0, 31)`` and is incremented before each ``update()``
(jobshoplab/types/stochasticy_models.py:9-30).  The n-th value of a time object is therefore a pure function
of (distribution, start seed + n): the first variate of ``default_rng(start + n)``.  This module tabulates
that function on the host with numpy itself (the third-party dependency that defines it; numpy >= 2.3 per
the reference's pyproject.toml) for seeds 0..depth-1, so the device reproduces the reference's actual --
low-entropy -- sample distribution exactly, and bit-exactly when given the same start seeds.
Runs once per instance at lowering time; nothing here is on the step path.
"""
from __future__ import annotations

import numpy as np

from . import lowered as L

DEPTH = 256  # seeds 0..255: start seed (0..30) + up to 225 updates per object per episode


def _value(kind: int, base: int, param: float, seed: int) -> int:
    rng = np.random.default_rng(seed=seed)
    if kind == L.TIME_POISSON:
        v = base + rng.poisson(base + 0.5)                       # stochasticy_models.py:44-51
    elif kind == L.TIME_UNIFORM:
        v = int(rng.uniform(base - param, base + param))         # :64-71
    elif kind == L.TIME_GAUSSIAN:
        v = int(rng.normal(base, param))                         # :87-93
    elif kind == L.TIME_GAMMA:
        v = int(round(base + rng.gamma(base, param)))            # :107-121 (shape = base_time, sic)
    else:
        raise ValueError(kind)
    return max(0, int(v))                                        # :15, :30


def flat_specs(li: L.LoweredInstance):
    """(kind, base, param) of every time object in the engine's flat order: ops, setup, travel,
    machine-outage frequency, duration, transport-outage frequency, duration."""
    fams = ((li.op_duration, li.op_spec), (li.setup_time, li.setup_spec), (li.travel_time, li.travel_spec),
            (li.m_outage_freq, li.m_ofreq_spec), (li.m_outage_dur, li.m_odur_spec),
            (li.t_outage_freq, li.t_ofreq_spec), (li.t_outage_dur, li.t_odur_spec))
    kind, base, param = [], [], []
    for time, sp in fams:
        n = len(time)
        if sp.kind is None:
            kind.append(np.zeros(n, np.int32)); base.append(np.asarray(time, np.int32)); param.append(np.zeros(n, np.float32))
        else:
            kind.append(sp.kind); base.append(sp.base); param.append(sp.param)
    return np.concatenate(kind), np.concatenate(base), np.concatenate(param)


def sample_tables(li: L.LoweredInstance, depth: int = DEPTH):
    """-> (table int32 [rows][depth], row int32 [n_tobj] (-1 = static))."""
    kind, base, param = flat_specs(li)
    rows, index = {}, np.full(len(kind), -1, dtype=np.int32)
    for i in np.nonzero(kind != L.TIME_STATIC)[0]:
        key = (int(kind[i]), int(base[i]), float(param[i]))
        if key not in rows:
            rows[key] = (len(rows), [_value(*key, s) for s in range(depth)])
        index[i] = rows[key][0]
    table = np.zeros((max(len(rows), 1), depth), dtype=np.int32)
    for r, vals in rows.values():
        table[r] = vals
    return table, index
