"""LoweredInstance: the structure-of-arrays form of a compiled JobShopLab instance.

This is what the reference's ``Compiler.compile()`` (jobshoplab/compiler/compiler.py:84-106)
hands to the state machine -- ``(InstanceConfig, State)`` -- flattened to int32 arrays in the
index spaces of ``include/jobshoplab_b200.h``:

* buffer idx: pre(m)=3m, post(m)=3m+1, machine buffer(m)=3m+2, agv(t)=3M+t, standalone s=3M+T+s
* place     : machine m -> m, standalone buffer s -> M+s (axis of the travel matrix)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BUF_FIFO, BUF_LIFO, BUF_DUMMY, BUF_FLEX = 0, 1, 2, 3
ROLE_INPUT, ROLE_OUTPUT, ROLE_COMPONENT, ROLE_COMPENSATION = 0, 1, 2, 3
TIME_STATIC, TIME_POISSON, TIME_UNIFORM, TIME_GAUSSIAN, TIME_GAMMA = 0, 1, 2, 3, 4
CAP_INF = 0x7FFFFFFF
NO_TIME = -1


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


@dataclass
class TimeSpec:
    """Stochastic behaviour of one family of time objects (None = static)."""

    kind: np.ndarray | None = None
    base: np.ndarray | None = None
    param: np.ndarray | None = None

    @property
    def is_static(self) -> bool:
        return self.kind is None or not np.any(self.kind != TIME_STATIC)


@dataclass
class LoweredInstance:
    description: str
    n_jobs: int
    n_machines: int
    n_transports: int
    n_sbuf: int
    n_tools: int
    job_op_start: np.ndarray
    op_machine: np.ndarray
    op_tool: np.ndarray
    op_duration: np.ndarray
    pre_type: np.ndarray
    pre_cap: np.ndarray
    post_type: np.ndarray
    post_cap: np.ndarray
    setup_time: np.ndarray
    default_tool: int
    sbuf_type: np.ndarray
    sbuf_cap: np.ndarray
    sbuf_role: np.ndarray
    buf_ref_id: np.ndarray
    travel_time: np.ndarray
    agv_init_place: np.ndarray
    job_init_buf: np.ndarray
    start_time: int
    m_outage_start: np.ndarray
    m_outage_freq: np.ndarray
    m_outage_dur: np.ndarray
    t_outage_freq: np.ndarray
    t_outage_dur: np.ndarray
    lower_bound: int
    max_allowed_time: int
    tools: list = field(default_factory=list)
    op_spec: TimeSpec = field(default_factory=TimeSpec)
    setup_spec: TimeSpec = field(default_factory=TimeSpec)
    travel_spec: TimeSpec = field(default_factory=TimeSpec)
    m_ofreq_spec: TimeSpec = field(default_factory=TimeSpec)
    m_odur_spec: TimeSpec = field(default_factory=TimeSpec)
    t_ofreq_spec: TimeSpec = field(default_factory=TimeSpec)
    t_odur_spec: TimeSpec = field(default_factory=TimeSpec)

    def __post_init__(self):
        for name in ("job_op_start", "op_machine", "op_tool", "op_duration", "pre_type", "pre_cap",
                     "post_type", "post_cap", "setup_time", "sbuf_type", "sbuf_cap", "sbuf_role",
                     "buf_ref_id", "travel_time", "agv_init_place", "job_init_buf", "m_outage_start",
                     "m_outage_freq", "m_outage_dur", "t_outage_freq", "t_outage_dur"):
            setattr(self, name, _i32(getattr(self, name)))
        for sp in self.specs():
            if sp.kind is not None:
                sp.kind, sp.base, sp.param = _i32(sp.kind), _i32(sp.base), _f32(sp.param)

    def specs(self):
        return (self.op_spec, self.setup_spec, self.travel_spec, self.m_ofreq_spec, self.m_odur_spec,
                self.t_ofreq_spec, self.t_odur_spec)

    @property
    def n_ops(self) -> int:
        return int(self.job_op_start[-1])

    @property
    def n_buffers(self) -> int:
        return 3 * self.n_machines + self.n_transports + self.n_sbuf

    @property
    def n_places(self) -> int:
        return self.n_machines + self.n_sbuf

    @property
    def is_deterministic(self) -> bool:
        return all(sp.is_static for sp in self.specs())

    def same_shape(self, other: "LoweredInstance") -> bool:
        a = (self.n_jobs, self.n_machines, self.n_transports, self.n_sbuf, self.n_tools, self.n_ops,
             len(self.m_outage_freq), len(self.t_outage_freq))
        b = (other.n_jobs, other.n_machines, other.n_transports, other.n_sbuf, other.n_tools, other.n_ops,
             len(other.m_outage_freq), len(other.t_outage_freq))
        return a == b and np.array_equal(self.job_op_start, other.job_op_start)
