"""Lowering of the REFERENCE's own objects -- ``(InstanceConfig, State)`` as returned by
``jobshoplab.compiler.Compiler.compile()`` (compiler.py:84-106) -- into a ``LoweredInstance``.

This is what a plug-in that sits behind the reference's middleware slot (SURVEY 8(b) seam 2) receives: the reference
env has already parsed / validated / manipulated the instance (env.py:491-496) and hands the resulting dataclasses to
``Middleware.__init__(instance=...)`` and ``Middleware.reset(init_state)``.  Nothing here imports the reference: the
objects are read by attribute and their classes are told apart by name, so the module loads without it (the
reference only has to be importable where its objects come from).

Index spaces and field meanings: include/jobshoplab_b200.h (jsl_instance_desc); id helpers follow
utils/utils.py:104-170 (get_id_int & friends).
"""
from __future__ import annotations

import numpy as np

from . import lowered as L

_BUF_TYPE = {"FIFO": L.BUF_FIFO, "LIFO": L.BUF_LIFO, "DUMMY": L.BUF_DUMMY, "FLEX_BUFFER": L.BUF_FLEX}
_BUF_ROLE = {"INPUT": L.ROLE_INPUT, "OUTPUT": L.ROLE_OUTPUT, "COMPONENT": L.ROLE_COMPONENT,
             "COMPENSATION": L.ROLE_COMPENSATION}
_TIME_KIND = {"PoissonFunction": L.TIME_POISSON, "UniformFunction": L.TIME_UNIFORM,
              "GaussianFunction": L.TIME_GAUSSIAN, "GammaFunction": L.TIME_GAMMA}


def tool_list(instance) -> list[str]:
    """Tools the state machine can ever mount: "tl-0" (mapper.py:396) plus every operation's tool, sorted."""
    tools = {"tl-0"}
    for job in instance.instance.specification:
        for op in job.operations:
            tools.add(op.tool)
    return sorted(tools)


class Indexer:
    """String ids of the reference <-> the engine's index spaces."""

    def __init__(self, instance):
        self.instance = instance
        self.jobs = [j.id for j in instance.instance.specification]
        self.machines = [m.id for m in instance.machines]
        self.transports = [t.id for t in instance.transports]
        self.sbufs = [b.id for b in instance.buffers]
        self.J, self.M, self.T, self.S = len(self.jobs), len(self.machines), len(self.transports), len(self.sbufs)
        self.job = {k: i for i, k in enumerate(self.jobs)}
        self.machine = {k: i for i, k in enumerate(self.machines)}
        self.transport = {k: i for i, k in enumerate(self.transports)}
        self.buffer, self.buffer_ids = {}, [None] * (3 * self.M + self.T + self.S)
        for i, m in enumerate(instance.machines):
            for off, b in ((0, m.prebuffer), (1, m.postbuffer), (2, m.buffer)):
                self.buffer[b.id] = 3 * i + off
        for t, tr in enumerate(instance.transports):
            self.buffer[tr.buffer.id] = 3 * self.M + t
        for s, b in enumerate(instance.buffers):
            self.buffer[b.id] = 3 * self.M + self.T + s
        for k, i in self.buffer.items():
            self.buffer_ids[i] = k
        self.sbuf = {k: s for s, k in enumerate(self.sbufs)}
        self.tools = tool_list(instance)
        self.tool = {t: i for i, t in enumerate(self.tools)}

    def place(self, s: str) -> int:
        return self.machine[s] if s in self.machine else self.M + self.sbuf[s]

    def place_id(self, p: int) -> str:
        return self.machines[p] if p < self.M else self.sbufs[p - self.M]


def _time_family(objs):
    """list of DeterministicTimeConfig | StochasticTimeConfig | None -> (time int32[], TimeSpec, start seeds)."""
    n = len(objs)
    time = np.full(n, L.NO_TIME, dtype=np.int32)
    kind, base = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    param = np.zeros(n, dtype=np.float32)
    seeds = np.full(n, -1, dtype=np.int32)
    for i, o in enumerate(objs):
        if o is None:
            continue
        time[i] = base[i] = int(o.time)
        k = _TIME_KIND.get(type(o).__name__)
        if k is None:
            continue  # DeterministicTimeConfig
        kind[i], base[i] = k, int(o.base_time)
        seeds[i] = int(getattr(o, "_current_seed", -1))
        if k == L.TIME_UNIFORM:
            param[i] = o.high - o.base_time
        elif k == L.TIME_GAUSSIAN:
            param[i] = o.std
        elif k == L.TIME_GAMMA:
            param[i] = o.scale
    return time, (L.TimeSpec(kind, base, param) if kind.any() else L.TimeSpec()), seeds


def _taillard_lower_bound(instance, job_op_start, op_dur) -> int:
    """utils/utils.py:310-331 calculate_lower_bound on the reference objects: same routine as the product compiler
    (machine numbers are the numeric part of the machine ids, as the reference takes them)."""
    from .compiler import _taillard_lower_bound as lb

    mach = [int(o.machine.split("-")[1]) for job in instance.instance.specification for o in job.operations]
    return int(lb(np.asarray(job_op_start), np.asarray(mach), np.asarray(op_dur)))


def lower_reference(instance, init_state, ix: Indexer | None = None) -> L.LoweredInstance:
    """(InstanceConfig, State) of the reference -> LoweredInstance.  ``.start_seeds`` carries the numpy start seeds
    of the stochastic time objects (engine flat order: ops | setup | travel | machine-outage freq | dur |
    transport-outage freq | dur) so that jsl_batch_set_start_seeds reproduces the reference's own sample streams."""
    ix = ix or Indexer(instance)
    M, T, S = ix.M, ix.T, ix.S
    tools = ix.tools
    cap = lambda c: L.CAP_INF if c >= L.CAP_INF else int(c)
    btype = lambda b: _BUF_TYPE[b.type.name]
    job_op_start, ops = [0], []
    for job in instance.instance.specification:
        ops += list(job.operations)
        job_op_start.append(len(ops))
    op_dur, op_spec, s0 = _time_family([o.duration for o in ops])
    setup_time, setup_spec, s1 = _time_family([m.setup_times.get((a, b)) for m in instance.machines for a in tools for b in tools])
    places = ix.machines + ix.sbufs
    travel_time, travel_spec, s2 = _time_family([instance.logistics.travel_times.get((a, b)) for a in places for b in places])
    m_out_start, m_out = [0], []
    for m in instance.machines:
        m_out += list(m.outages)
        m_out_start.append(len(m_out))
    m_ofreq, m_ofreq_spec, s3 = _time_family([o.frequency for o in m_out])
    m_odur, m_odur_spec, s4 = _time_family([o.duration for o in m_out])
    t_out = list(instance.transports[0].outages) if T else []
    for tr in instance.transports:  # mapper.py:1082-1092: one shared outage list for all AGVs
        if [o.id for o in tr.outages] != [o.id for o in t_out]:
            raise ValueError("transports with different outage lists are not supported")
    t_ofreq, t_ofreq_spec, s5 = _time_family([o.frequency for o in t_out])
    t_odur, t_odur_spec, s6 = _time_family([o.duration for o in t_out])
    buf_ref = np.zeros(3 * M + T + S, dtype=np.int32)
    for bid, bi in ix.buffer.items():
        buf_ref[bi] = int(bid.split("-")[1])
    agv_place = []
    for t in init_state.transports:
        loc = t.location.location
        if not isinstance(loc, str):
            raise ValueError("initial transport locations must be places")
        agv_place.append(ix.place(loc))
    li = L.LoweredInstance(
        description=str(instance.description), n_jobs=ix.J, n_machines=M, n_transports=T, n_sbuf=S, n_tools=len(tools),
        job_op_start=job_op_start, op_machine=[ix.machine[o.machine] for o in ops],
        op_tool=[ix.tool[o.tool] for o in ops], op_duration=op_dur,
        pre_type=[btype(m.prebuffer) for m in instance.machines], pre_cap=[cap(m.prebuffer.capacity) for m in instance.machines],
        post_type=[btype(m.postbuffer) for m in instance.machines], post_cap=[cap(m.postbuffer.capacity) for m in instance.machines],
        setup_time=setup_time, default_tool=ix.tool["tl-0"],
        sbuf_type=[btype(b) for b in instance.buffers], sbuf_cap=[cap(b.capacity) for b in instance.buffers],
        sbuf_role=[_BUF_ROLE[b.role.name] for b in instance.buffers], buf_ref_id=buf_ref, travel_time=travel_time,
        agv_init_place=agv_place, job_init_buf=[ix.buffer[j.location] for j in init_state.jobs],
        start_time=int(init_state.time.time), m_outage_start=m_out_start, m_outage_freq=m_ofreq, m_outage_dur=m_odur,
        t_outage_freq=t_ofreq, t_outage_dur=t_odur,
        lower_bound=_taillard_lower_bound(instance, job_op_start, op_dur), max_allowed_time=int(sum(int(o.duration.time) for o in ops)),
        tools=tools, op_spec=op_spec, setup_spec=setup_spec, travel_spec=travel_spec, m_ofreq_spec=m_ofreq_spec,
        m_odur_spec=m_odur_spec, t_ofreq_spec=t_ofreq_spec, t_odur_spec=t_odur_spec)
    li.start_seeds = np.concatenate([s0, s1, s2, s3, s4, s5, s6]).astype(np.int32)
    return li
