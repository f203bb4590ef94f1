"""Reference harness -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Runs the UNMODIFIED reference at /root/reference (behind the import shims in
oracle/shims) in THIS container and turns its nested-dataclass state into the
canonical int32 digest that the C oracle (oracle/jsl_oracle.c) and the CUDA
engine (jsl_get_state) also emit.  Only oracle/gen_golden.py and ad-hoc
validation scripts use it; /root/reference does not exist on the GPU box.

Digest layout (all int32; see DESIGN.md "canonical digest"):
  [time, J, M, T, S]
  per job:        n_done, running, loc(buffer idx), then per op: state(0 idle/1 proc/2 done), start, end (-1 = NoTime)
  per machine:    state(0 I/1 S/2 W/3 O), occupied_till(-1 NoTime), tool idx,
                  len(pre), pre..., len(buf), buf..., len(post), post...
  per transport:  state(0 Idle/1 Working/2 Pickup/3 Transit/4 Outage/5 WaitingPickup),
                  occ_kind(0 NoTime/1 Time/2 TimeDependency), occ(time | blocking job),
                  td_buffer, td_tr_comp, td_tr_job, carried job, transport_job,
                  loc_kind(0 place/1 triple), loc0(place), loc1(buffer idx), loc2(place)
  per standalone buffer: len(store), store...
  per machine outage, then per transport outage: active(0/1), a(start | last_active), b(end | -1)
  n_possible, then per possible transition: kind(0 machine/1 transport), comp idx, new_state code, job
Buffer idx: pre(m)=3m, post(m)=3m+1, machine buffer(m)=3m+2, agv(t)=3M+t, standalone s=3M+T+s.
Place: machine m -> m ; standalone buffer s -> M+s.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
def _ref_root() -> Path:
    """/root/reference where it exists (the build container), else the installed copy oracle/_ref (oracle/make_ref.py;
    it travels to the GPU box with the repo snapshot)."""
    env = os.environ.get("JSL_REFERENCE_ROOT")
    for p in ([Path(env)] if env else []) + [Path("/root/reference"), _HERE / "_ref"]:
        if (p / "jobshoplab" / "__init__.py").exists():
            return p
    return Path("/root/reference")


REF_ROOT = _ref_root()


def _activate():
    shims = str(_HERE / "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(1, str(REF_ROOT))


_activate()

import logging  # noqa: E402

logging.disable(logging.CRITICAL)

from jobshoplab import JobShopLabEnv  # noqa: E402
from jobshoplab.compiler import Compiler  # noqa: E402
from jobshoplab.compiler.repos import DslRepository, DslStrRepository, SpecRepository  # noqa: E402
from jobshoplab.types.config_types import Config  # noqa: E402
from jobshoplab.types.state_types import (  # noqa: E402
    MachineStateState,
    NoTime,
    OperationStateState,
    OutageActive,
    Time,
    TimeDependency,
    TransportStateState,
)
from jobshoplab.utils.state_machine_utils import dispatch_rules_uitls as dispatch_utils  # noqa: E402

M_STATE = {
    MachineStateState.IDLE: 0,
    MachineStateState.SETUP: 1,
    MachineStateState.WORKING: 2,
    MachineStateState.OUTAGE: 3,
}
T_STATE = {
    TransportStateState.IDLE: 0,
    TransportStateState.WORKING: 1,
    TransportStateState.PICKUP: 2,
    TransportStateState.TRANSIT: 3,
    TransportStateState.OUTAGE: 4,
    TransportStateState.WAITINGPICKUP: 5,
}
O_STATE = {
    OperationStateState.IDLE: 0,
    OperationStateState.PROCESSING: 1,
    OperationStateState.DONE: 2,
}


def make_config(**over) -> Config:
    """Default `Config()` with a few dotted overrides, e.g.
    allow_early_transport=False, truncation_active=True, truncation_joker=3."""
    cfg = Config()
    if "allow_early_transport" in over:
        cfg.state_machine.allow_early_transport = bool(over["allow_early_transport"])
    mw = cfg.middleware.event_based_binary_action_middleware
    if "truncation_active" in over:
        mw.truncation_active = bool(over["truncation_active"])
    if "truncation_joker" in over:
        mw.truncation_joker = int(over["truncation_joker"])
    rw = cfg.reward_factory.binary_action_jssp_reward
    for k in ("sparse_bias", "dense_bias", "truncation_bias"):
        if k in over:
            setattr(rw, k, float(over[k]))
    if "observation_factory" in over:
        cfg.env.observation_factory = over["observation_factory"]
    return cfg


class RecordingCompiler:
    """Wraps the reference Compiler: lowers (instance, init_state) the moment they are compiled --
    before the env's reset step mutates any stochastic time object -- and keeps them."""

    def __init__(self, inner):
        self.inner = inner
        self.lowered = None
        self.instance = None
        self.init_state = None

    def compile(self):
        inst, init = self.inner.compile()
        self.instance, self.init_state = inst, init
        SAMPLE_TAPE.clear()
        self.lowered = lower_reference(inst, init)
        return inst, init


SAMPLE_TAPE: list[int] = []  # every StochasticTimeConfig.update() result, in call order


def _install_sample_recorder():
    from jobshoplab.types.stochasticy_models import StochasticTimeConfig

    if getattr(StochasticTimeConfig, "_jsl_recorded", False):
        return
    orig = StochasticTimeConfig.update

    def update(self):
        orig(self)
        SAMPLE_TAPE.append(int(self.time))

    StochasticTimeConfig.update = update
    StochasticTimeConfig._jsl_recorded = True


def make_env(kind: str, src, seed=None, **cfg_over) -> JobShopLabEnv:
    """kind: 'spec' (Taillard spec file path), 'dsl' (yaml path), 'dslstr' (yaml text / dict).
    The returned env carries `.jsl_rc` (RecordingCompiler with `.lowered`)."""
    _install_sample_recorder()
    cfg = make_config(**cfg_over)
    if kind == "spec":
        repo = SpecRepository(Path(src), loglevel="warning", config=cfg)
    elif kind == "dsl":
        repo = DslRepository(Path(src), loglevel="warning", config=cfg)
    elif kind == "dslstr":
        if not isinstance(src, str):
            import yaml
            src = yaml.safe_dump(src)
        repo = DslStrRepository(src, loglevel="warning", config=cfg)
    else:
        raise ValueError(kind)
    rc = RecordingCompiler(Compiler(cfg, loglevel="warning", repo=repo))
    env = JobShopLabEnv(config=cfg, compiler=rc, seed=seed)
    env.jsl_rc = rc
    return env


def tool_list(instance) -> list[str]:
    """Tools the state machine can ever mount: "tl-0" plus every operation's tool, sorted."""
    tools = {"tl-0"}
    for job in instance.instance.specification:
        for op in job.operations:
            tools.add(op.tool)
    return sorted(tools)


class Indexer:
    def __init__(self, instance):
        self.instance = instance
        self.J = len(instance.instance.specification)
        self.M = len(instance.machines)
        self.T = len(instance.transports)
        self.S = len(instance.buffers)
        self.job = {j.id: i for i, j in enumerate(instance.instance.specification)}
        self.machine = {m.id: i for i, m in enumerate(instance.machines)}
        self.transport = {t.id: i for i, t in enumerate(instance.transports)}
        self.buffer = {}
        for i, m in enumerate(instance.machines):
            self.buffer[m.prebuffer.id] = 3 * i
            self.buffer[m.postbuffer.id] = 3 * i + 1
            self.buffer[m.buffer.id] = 3 * i + 2
        for t, tr in enumerate(instance.transports):
            self.buffer[tr.buffer.id] = 3 * self.M + t
        for s, b in enumerate(instance.buffers):
            self.buffer[b.id] = 3 * self.M + self.T + s
        self.sbuf = {b.id: s for s, b in enumerate(instance.buffers)}
        self.tools = {t: i for i, t in enumerate(tool_list(instance))}

    def place(self, s: str) -> int:
        if s in self.machine:
            return self.machine[s]
        if s in self.sbuf:
            return self.M + self.sbuf[s]
        raise KeyError(s)

    def tcode(self, t) -> int:
        if isinstance(t, Time):
            return int(t.time)
        return -1

    def transition(self, tr):
        if tr.component_id in self.machine:
            return (0, self.machine[tr.component_id], M_STATE[tr.new_state],
                    self.job[tr.job_id] if tr.job_id is not None else -1)
        return (1, self.transport[tr.component_id], T_STATE[tr.new_state],
                self.job[tr.job_id] if tr.job_id is not None else -1)


def digest(ix: Indexer, result) -> np.ndarray:
    """Canonical digest of a StateMachineResult (state + possible transitions)."""
    st = result.state
    out = [ix.tcode(st.time), ix.J, ix.M, ix.T, ix.S]
    for job in st.jobs:
        n_done = sum(op.operation_state_state == OperationStateState.DONE for op in job.operations)
        running = any(op.operation_state_state == OperationStateState.PROCESSING for op in job.operations)
        out += [n_done, int(running), ix.buffer[job.location]]
        for op in job.operations:
            out += [O_STATE[op.operation_state_state], ix.tcode(op.start_time), ix.tcode(op.end_time)]
    for m in st.machines:
        out += [M_STATE[m.state], ix.tcode(m.occupied_till), ix.tools[m.mounted_tool]]
        for buf in (m.prebuffer, m.buffer, m.postbuffer):
            out.append(len(buf.store))
            out += [ix.job[j] for j in buf.store]
    for t in st.transports:
        occ = t.occupied_till
        if isinstance(occ, TimeDependency):
            tr = occ.transition
            o = [2, ix.job[occ.job_id], ix.buffer[occ.buffer_id], ix.transport[tr.component_id],
                 ix.job[tr.job_id] if tr.job_id is not None else -1]
        elif isinstance(occ, Time):
            o = [1, int(occ.time), -1, -1, -1]
        else:
            o = [0, -1, -1, -1, -1]
        carried = ix.job[t.buffer.store[0]] if len(t.buffer.store) else -1
        tj = ix.job[t.transport_job] if t.transport_job is not None else -1
        loc = t.location.location
        if isinstance(loc, str):
            l = [0, ix.place(loc), -1, -1]
        else:
            l = [1, ix.place(loc[0]), ix.buffer[loc[1]], ix.place(loc[2])]
        out += [T_STATE[t.state]] + o + [carried, tj] + l
    for b in st.buffers:
        out.append(len(b.store))
        out += [ix.job[j] for j in b.store]
    for comp in tuple(st.machines) + tuple(st.transports):
        for o in comp.outages:
            if isinstance(o.active, OutageActive):
                out += [1, ix.tcode(o.active.start_time), ix.tcode(o.active.end_time)]
            else:
                out += [0, ix.tcode(o.active.last_time_active), -1]
    out.append(len(result.possible_transitions))
    for tr in result.possible_transitions:
        out += list(ix.transition(tr))
    return np.asarray(out, dtype=np.int32)


def fnv1a64(buf: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in buf:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def fnv1a64_np(a: np.ndarray) -> int:
    return fnv1a64(np.ascontiguousarray(a).tobytes())


def obs_bytes(obs: dict) -> bytes:
    """Packed default observation (BinaryActionObservationFactory), the engine's record layout:
    f32 current_time | f32 current_transition[3] | i32 job_progression[J] | i32 machine_progression[M]
    | i8 job_running[J] | i8 available_jobs[J] | i8 machine_running[M] | i8 job_executed_on_machine[J*M],
    zero-padded to a multiple of 16 bytes."""
    parts = [
        np.asarray(obs["current_time"], dtype=np.float32).reshape(-1),
        np.asarray(obs["current_transition"], dtype=np.float32).reshape(-1),
        np.asarray(obs["job_progression"], dtype=np.int32),
        np.asarray(obs["machine_progression"], dtype=np.int32),
        np.asarray(obs["job_running"], dtype=np.int8),
        np.asarray(obs["available_jobs"], dtype=np.int8),
        np.asarray(obs["machine_running"], dtype=np.int8),
        np.asarray(obs["job_executed_on_machine"], dtype=np.int8).reshape(-1),
    ]
    raw = b"".join(p.tobytes() for p in parts)
    return raw + b"\0" * (-len(raw) % 16)


def obs_oparray_bytes(obs: dict) -> bytes:
    """Packed BinaryOperationArrayObservation: f32 current_transition[3] | f32 0 | f32 operation_state[N]
    | f32 job_locations[J], padded to 16 bytes."""
    parts = [
        np.asarray(obs["current_transition"], dtype=np.float32).reshape(-1),
        np.zeros(1, dtype=np.float32),
        np.asarray(obs["operation_state"], dtype=np.float32).reshape(-1),
        np.asarray(obs["job_locations"], dtype=np.float32).reshape(-1),
    ]
    raw = b"".join(p.tobytes() for p in parts)
    return raw + b"\0" * (-len(raw) % 16)


def obs_tassel_bytes(obs: dict) -> bytes:
    """Packed TasselJsspObservation: f32 [7][J] in the reference's key order, padded to 16 bytes."""
    keys = ("allocatable", "left_over_time", "percent_finished", "total_completion",
            "time_until_next_machine_is_free", "idle_since_last_op", "cum_idle_time")
    raw = b"".join(np.asarray(obs[k], dtype=np.float64).astype(np.float32).tobytes() for k in keys)
    return raw + b"\0" * (-len(raw) % 16)


# ---------------------------------------------------------------- policies
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011). counter: 4 u32, key: 2 u32 -> 4 u32."""
    M0, M1 = 0xD2511F53, 0xCD9E8D57
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def random_action(seed: int, instance_idx: int, step: int) -> int:
    """The engine's RANDOM policy: bit 0 of Philox4x32-10(counter=(step, 0, inst_lo, inst_hi),
    key=(seed_lo, seed_hi))[0]."""
    r = philox4x32_10((step, 0, instance_idx & 0xFFFFFFFF, instance_idx >> 32),
                      (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return r[0] & 1


def rule_action(env, rule: str) -> int:
    """fifo / spt / mwkr exactly as jobshoplab/utils/dipatching_benchmark.py:17-67."""
    tr = env.state.possible_transitions[0]
    if rule == "fifo" or tr.component_id.startswith("t"):
        return 1
    if rule == "spt":
        sel = dispatch_utils.get_machine_prebuffer_jobs_by_processing_time(
            env.instance, env.state.state, mode="SPT")
    elif rule == "mwkr":
        sel = dispatch_utils.get_machine_jobs_by_remaining_processing_time(
            env.instance, env.state.state, mode="srpt")
    else:
        raise ValueError(rule)
    return int(sel[tr.component_id] == tr.job_id)


# ---------------------------------------------------------------- lowering of reference objects
def lower_reference(instance, init_state):
    """(InstanceConfig, State) of the reference -> jobshoplab_b200.lowered.LoweredInstance.
    Used to validate the product's own compiler (jobshoplab_b200/compiler.py) and to feed the
    oracle the exact objects the reference simulates."""
    sys.path.insert(0, str(_HERE.parent))
    from jobshoplab_b200 import lowered as L
    from jobshoplab.types.instance_config_types import (BufferRoleConfig, BufferTypeConfig,
                                                         DeterministicTimeConfig)
    from jobshoplab.types.stochasticy_models import (GammaFunction, GaussianFunction, PoissonFunction,
                                                     UniformFunction)
    from jobshoplab.utils.utils import calculate_lower_bound, get_max_allowed_time

    ix = Indexer(instance)
    J, M, T, S = ix.J, ix.M, ix.T, ix.S
    tools = tool_list(instance)
    nt = len(tools)
    BT = {BufferTypeConfig.FIFO: L.BUF_FIFO, BufferTypeConfig.LIFO: L.BUF_LIFO,
          BufferTypeConfig.DUMMY: L.BUF_DUMMY, BufferTypeConfig.FLEX_BUFFER: L.BUF_FLEX}
    BR = {BufferRoleConfig.INPUT: L.ROLE_INPUT, BufferRoleConfig.OUTPUT: L.ROLE_OUTPUT,
          BufferRoleConfig.COMPONENT: L.ROLE_COMPONENT, BufferRoleConfig.COMPENSATION: L.ROLE_COMPENSATION}

    def cap(c):
        return L.CAP_INF if c >= L.CAP_INF else int(c)

    seeds_all = []

    def tspec(objs):
        """list of time objects (or None) -> (time array, TimeSpec)"""
        seeds_all.extend(-1 if (o is None or isinstance(o, DeterministicTimeConfig)) else int(o._current_seed)
                         for o in objs)
        time = np.full(len(objs), L.NO_TIME, dtype=np.int32)
        kind = np.zeros(len(objs), dtype=np.int32)
        base = np.zeros(len(objs), dtype=np.int32)
        param = np.zeros(len(objs), dtype=np.float32)
        for i, o in enumerate(objs):
            if o is None:
                continue
            time[i] = o.time
            base[i] = o.time
            if isinstance(o, DeterministicTimeConfig):
                continue
            base[i] = o.base_time
            if isinstance(o, PoissonFunction):
                kind[i] = L.TIME_POISSON
            elif isinstance(o, UniformFunction):
                kind[i], param[i] = L.TIME_UNIFORM, o.high - o.base_time
            elif isinstance(o, GaussianFunction):
                kind[i], param[i] = L.TIME_GAUSSIAN, o.std
            elif isinstance(o, GammaFunction):
                kind[i], param[i] = L.TIME_GAMMA, o.scale
            else:
                raise TypeError(o)
        sp = L.TimeSpec(kind, base, param) if kind.any() else L.TimeSpec()
        return time, sp

    job_op_start = [0]
    ops = []
    for job in instance.instance.specification:
        ops += list(job.operations)
        job_op_start.append(len(ops))
    op_dur, op_spec = tspec([o.duration for o in ops])
    setup_objs = []
    for m in instance.machines:
        for a in tools:
            for b in tools:
                setup_objs.append(m.setup_times.get((a, b)))
    setup_time, setup_spec = tspec(setup_objs)
    places = [m.id for m in instance.machines] + [b.id for b in instance.buffers]
    travel_objs = [instance.logistics.travel_times.get((a, b)) for a in places for b in places]
    travel_time, travel_spec = tspec(travel_objs)
    m_out_start = [0]
    m_out = []
    for m in instance.machines:
        m_out += list(m.outages)
        m_out_start.append(len(m_out))
    m_ofreq, m_ofreq_spec = tspec([o.frequency for o in m_out])
    m_odur, m_odur_spec = tspec([o.duration for o in m_out])
    t_out = list(instance.transports[0].outages) if T else []
    for tr in instance.transports:
        assert [o.id for o in tr.outages] == [o.id for o in t_out]
    t_ofreq, t_ofreq_spec = tspec([o.frequency for o in t_out])
    t_odur, t_odur_spec = tspec([o.duration for o in t_out])
    buf_ref = np.zeros(3 * M + T + S, dtype=np.int32)
    for bid, bi in ix.buffer.items():
        buf_ref[bi] = int(bid.split("-")[1])
    agv_place = []
    for t in init_state.transports:
        assert isinstance(t.location.location, str)
        agv_place.append(ix.place(t.location.location))
    li = L.LoweredInstance(
        description=str(instance.description),
        n_jobs=J, n_machines=M, n_transports=T, n_sbuf=S, n_tools=nt,
        job_op_start=job_op_start,
        op_machine=[ix.machine[o.machine] for o in ops],
        op_tool=[tools.index(o.tool) for o in ops],
        op_duration=op_dur,
        pre_type=[BT[m.prebuffer.type] for m in instance.machines],
        pre_cap=[cap(m.prebuffer.capacity) for m in instance.machines],
        post_type=[BT[m.postbuffer.type] for m in instance.machines],
        post_cap=[cap(m.postbuffer.capacity) for m in instance.machines],
        setup_time=setup_time, default_tool=tools.index("tl-0"),
        sbuf_type=[BT[b.type] for b in instance.buffers],
        sbuf_cap=[cap(b.capacity) for b in instance.buffers],
        sbuf_role=[BR[b.role] for b in instance.buffers],
        buf_ref_id=buf_ref, travel_time=travel_time,
        agv_init_place=agv_place,
        job_init_buf=[ix.buffer[j.location] for j in init_state.jobs],
        start_time=int(init_state.time.time),
        m_outage_start=m_out_start, m_outage_freq=m_ofreq, m_outage_dur=m_odur,
        t_outage_freq=t_ofreq, t_outage_dur=t_odur,
        lower_bound=int(calculate_lower_bound(instance)),
        max_allowed_time=int(get_max_allowed_time(instance)),
        tools=tools,
        op_spec=op_spec, setup_spec=setup_spec, travel_spec=travel_spec,
        m_ofreq_spec=m_ofreq_spec, m_odur_spec=m_odur_spec,
        t_ofreq_spec=t_ofreq_spec, t_odur_spec=t_odur_spec,
    )
    # numpy start seeds of the stochastic objects, in the engine's flat time-object order (ops, setup, travel,
    # machine-outage freq, dur, transport-outage freq, dur) -- tspec() is called in exactly that order
    li.start_seeds = np.asarray(seeds_all, dtype=np.int32)
    return li
