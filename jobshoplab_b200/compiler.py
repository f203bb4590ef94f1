"""Instance compiler: Taillard spec files / the YAML DSL  ->  LoweredInstance (SoA int32).

Host-side mirror of the reference's ``jobshoplab.compiler`` for the env-step path: the same
repository classes and ``Compiler(config, loglevel, repo).compile()`` entry point
(jobshoplab/compiler/compiler.py:84-106, repos.py:78-270), and the same mapping rules and quirks as
``DictToInstanceMapper`` / ``DictToInitStateMapper`` (jobshoplab/compiler/mapper.py:1108-1407), but it
emits the flat arrays of ``jobshoplab_b200.lowered.LoweredInstance`` instead of nested dataclasses.
It runs once per instance on the host; nothing here is on the GPU step path.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Any

import numpy as np
import yaml

from . import lowered as L
from .exceptions import (InvalidSetupTimesError, InvalidToolUsageError, InvalidTransportConfig, InvalidValue,
                         MissingSpecificationError, NotImplementedError_, UnknownDistributionTypeError,
                         UnknownLocationNameError, InvalidDurationError, InvalidTimeBehaviorError,
                         InvalidOutageTypeError, ComponentAssociationError, InvalidDistributionError)


# --------------------------------------------------------------------------- repositories
class Repository:
    def load_as_dict(self) -> dict:
        raise NotImplementedError


class DslRepository(Repository):
    """YAML file repository (repos.py:78-128)."""

    def __init__(self, dir, loglevel=None, config=None, *a, **k):
        self.dir = Path(dir)
        if not self.dir.exists():
            raise FileNotFoundError(self.dir)

    def load_as_dict(self) -> dict:
        return yaml.safe_load(self.dir.read_text())

    def __repr__(self):
        return f"DslRepository(dir={self.dir})"


class DslStrRepository(Repository):
    """YAML string repository (repos.py:131-166)."""

    def __init__(self, dsl_str, loglevel=None, config=None, *a, **k):
        self.dsl_str = dsl_str

    def load_as_dict(self) -> dict:
        return yaml.safe_load(self.dsl_str)

    def __repr__(self):
        return "DslStrRepository(...)"


class SpecRepository(Repository):
    """Taillard-style spec file: first 2-token line "J M", then `machine duration ...` rows
    (repos.py:169-270)."""

    def __init__(self, dir, loglevel=None, config=None, *a, **k):
        self.dir = Path(dir)
        if not self.dir.exists():
            raise FileNotFoundError(self.dir)

    @staticmethod
    def text_to_dict(text: str) -> dict:
        lines = text.splitlines(keepends=True)
        first = next((i for i, ln in enumerate(lines) if len(ln.split()) == 2), None)
        if first is None:
            raise InvalidValue("Spec-File must contain 2 numbers in the first line", text)
        raw = "\n".join(lines[first + 1:])
        rows = raw.split("\n")
        n_m = int(len(rows[0].split(" ")) / 2)
        dsl = "|".join(f"(m{i},t)" for i in range(n_m)) + "\n"
        for line_num, line in enumerate(rows):
            nums = [int(x) for x in line.split()]
            dsl += f"j{line_num}|" + " ".join(f"({nums[i]}, {nums[i + 1]})" for i in range(0, len(nums), 2)) + "\n"
        return {"title": "InstanceConfig",
                "instance_config": {"description": "Minimal Spec-File Instance",
                                    "instance": {"specification": dsl,
                                                 "description": f"Specs from Spec-File {lines[first]}"}}}

    def load_as_dict(self) -> dict:
        return self.text_to_dict(self.dir.read_text())

    def __repr__(self):
        return f"SpecRepository(dir={self.dir})"


# --------------------------------------------------------------------------- mapping helpers
_JOB_LINE = re.compile(r"j\d+\|\(\d+,\d+\)+")
_INPUT_NAMES = ["input", "input-buffer", "inputbuffer", "input buffer", "input_buffer", "in-buf", "inbuf",
                "in buf", "in_buffer"]
_OUTPUT_NAMES = ["output", "output-buffer", "outputbuffer", "output buffer", "output_buffer", "out-buf"]
_BUF_TYPES = {"FIFO": L.BUF_FIFO, "LIFO": L.BUF_LIFO, "DUMMY": L.BUF_DUMMY, "FLEX_BUFFER": L.BUF_FLEX}
_BUF_ROLES = {"INPUT": L.ROLE_INPUT, "OUTPUT": L.ROLE_OUTPUT, "COMPONENT": L.ROLE_COMPONENT,
              "COMPENSATION": L.ROLE_COMPENSATION}


def _has_key(keys, d) -> bool:
    """mapper.py:471-487 (note: `in` on a str is a substring test -- kept on purpose)."""
    for k in keys:
        if k not in d:
            return False
        d = d[k]
    return True


class _IdCounter:
    """mapper.py:38-124."""

    def __init__(self):
        self.buffer_ids: list[str] = []

    def new_buffer_id(self) -> str:
        counter = len(self.buffer_ids)
        _id = f"b-{counter}"
        while _id in self.buffer_ids:
            _id = f"b-{counter}"
            counter += 1
        self.buffer_ids.append(_id)
        return _id


class _Time:
    """A DeterministicTimeConfig | StochasticTimeConfig: (time, kind, base, param)."""

    __slots__ = ("time", "kind", "base", "param")

    def __init__(self, time, kind=L.TIME_STATIC, base=None, param=0.0):
        self.time, self.kind, self.base, self.param = int(time), kind, int(time if base is None else base), float(param)


def _get_time(duration, time_behavior, sampler) -> _Time:
    """mapper.py:661-695 _get_time."""
    if duration is None:
        if isinstance(time_behavior, int) and not isinstance(time_behavior, bool):
            return _Time(time_behavior)
        if isinstance(time_behavior, dict) and "base" in time_behavior:
            duration = int(time_behavior["base"])
    if not isinstance(duration, int) or isinstance(duration, bool):
        raise InvalidDurationError(duration)
    if time_behavior == "static":
        return _Time(duration)
    if isinstance(time_behavior, dict):
        if "type" not in time_behavior:
            raise InvalidDistributionError("Distribution type must be specified")
        t = str(time_behavior["type"])
        if t == "poisson":
            kind, param = L.TIME_POISSON, 0.0
        elif t == "gamma":
            kind, param = L.TIME_GAMMA, float(time_behavior["scale"])
        elif t in ("uniform", "uni"):
            kind, param = L.TIME_UNIFORM, float(time_behavior["offset"])
        elif t in ("gaussian", "normal"):
            kind, param = L.TIME_GAUSSIAN, float(time_behavior["std"])
        else:
            raise UnknownDistributionTypeError(t)
        return _Time(sampler(kind, duration, param), kind, duration, param)
    raise InvalidTimeBehaviorError(time_behavior)


def _parse_matrix(text: str):
    """header|header...\\n row| v v v   (mapper.py:570-611, :795-820)."""
    lines = [ln.strip() for ln in text.strip().split("\n")]
    headers = lines[0].split("|")
    for ln in lines[1:]:
        parts = ln.split("|")
        row = parts[0]
        for col, v in zip(headers, map(int, parts[1].split())):
            yield row, col, v


def _buffer_spec(buf: dict, spec: dict) -> dict:
    """mapper.py:1010-1065 _add_buffer_spec."""
    for key, val in spec.items():
        if key == "capacity":
            buf["cap"] = val
        elif key == "type":
            s = val.upper()
            s = "FLEX_BUFFER" if s == "FLEX" else s
            if s not in _BUF_TYPES:
                raise InvalidTransportConfig(f"Unknown buffer type: {val}")
            buf["type"] = _BUF_TYPES[s]
        elif key == "name":
            pass
        elif key == "role":
            r = val.upper()
            if r not in _BUF_ROLES:
                raise InvalidValue("BufferRoleConfig", val)
            buf["role"] = _BUF_ROLES[r]
        elif key == "description":
            if not isinstance(val, str):
                raise InvalidValue("BufferDescription", val)
        else:
            raise InvalidValue("BufferTypeConfig", val)
    return buf


def _cap(c) -> int:
    return L.CAP_INF if c is None or c >= L.CAP_INF else int(c)


def _taillard_lower_bound(job_op_start, op_machine, op_duration) -> int:
    """utils.py:176-331 calculate_lower_bound (rectangular instances; K = ops per job)."""
    J = len(job_op_start) - 1
    ks = np.diff(job_op_start)
    if J == 0 or np.any(ks != ks[0]):
        raise InvalidValue("lower bound needs a rectangular instance", ks.tolist())
    K = int(ks[0])
    mach = np.asarray(op_machine).reshape(J, K)
    dur = np.asarray(op_duration, dtype=np.float64).reshape(J, K)
    best = 0.0
    for m in range(K):
        b = a = np.inf
        for j in range(J):
            hit = np.nonzero(mach[j] == m)[0]
            pos = int(hit[0]) if len(hit) else K
            b = min(b, dur[j, :pos].sum())
            op_id = pos if len(hit) else K - 1
            a = min(a, dur[j, op_id + 1:].sum())
        if np.any(mach >= K):
            raise InvalidValue("machine id out of range for the lower bound", int(mach.max()))
        T = dur[mach == m].sum()
        best = max(best, b + T + a)
    return int(max(best, dur.sum(axis=1).max()))


# --------------------------------------------------------------------------- the mapper
def lower_dict(spec_dict: dict, sampler=None) -> L.LoweredInstance:
    """DictToInstanceMapper.map + DictToInitStateMapper.map (mapper.py:1201, :1387) -> LoweredInstance.

    `sampler(kind, base, param) -> int` draws the construction-time sample of a stochastic time object
    (stochasticy_models.py:9-15); default: the base time (the engine redraws from its Philox stream)."""
    sampler = sampler or (lambda kind, base, param: max(0, int(base)))
    ic = spec_dict["instance_config"]
    # make_defaults (mapper.py:515-538)
    spec_lines = [ln.replace(" ", "") for ln in ic["instance"]["specification"].split("\n")]
    spec_lines = [ln for ln in spec_lines if _JOB_LINE.match(ln)]
    J = len(spec_lines)
    M = len(spec_lines[1].split(",")) - 1
    counter = _IdCounter()
    # standalone buffers given in the DSL are mapped first (mapper.py:1122-1130)
    sbufs: list[dict] = []
    if _has_key(("instance_config", "buffer"), spec_dict):
        for b in ic["buffer"]:
            buf = _buffer_spec({"id": b["name"], "type": L.BUF_FLEX, "cap": L.CAP_INF,
                                "role": L.ROLE_COMPENSATION}, b)
            if not buf["id"].startswith("b-"):
                raise InvalidValue("buffer_id", buf["id"])
            counter.buffer_ids.append(buf["id"])
            sbufs.append(buf)
    # machines (mapper.py:1131-1148, :822-850)
    pre, post, mbuf_ids, m_outages, setup_maps = [], [], [], [], []
    for m in range(M):
        mid = f"m-{m}"
        pb = {"id": counter.new_buffer_id(), "type": L.BUF_FLEX, "cap": L.CAP_INF, "role": L.ROLE_COMPONENT}
        qb = {"id": counter.new_buffer_id(), "type": L.BUF_FLEX, "cap": L.CAP_INF, "role": L.ROLE_COMPONENT}
        mbuf_ids.append(counter.new_buffer_id())
        outs = []
        if _has_key(("instance_config", "outages"), spec_dict):
            for o in ic["outages"]:
                if o["component"] in ["m", "machine", "Machine", "MACHINE", mid]:
                    outs.append(_outage(o, sampler))
        setup = None
        if _has_key(("instance_config", "setup_times"), spec_dict):
            st = next((i for i in ic["setup_times"] if i["machine"] == mid), None)
            if st is None:
                raise InvalidSetupTimesError(mid)
            tb = st["time_behavior"] if "time_behavior" in st.keys() else "static"
            setup = {(a, b): _get_time(v, tb, sampler) for a, b, v in _parse_matrix(st["specification"])}
        # global, then machine-specific buffer configs (mapper.py:852-987)
        for buf, key in ((pb, "prebuffer"), (qb, "postbuffer")):
            if _has_key(("instance_config", "machines", key), spec_dict):
                spec = next((s for s in ic["machines"][key] if s), None)
                if spec:
                    _buffer_spec(buf, spec)
        if _has_key(("instance_config", "machines"), spec_dict) and isinstance(ic["machines"], list):
            ms = next((s for s in ic["machines"] if isinstance(s, dict) and mid in s), None)
            mc = ms[mid] if ms else None
            if mc:
                for buf, key in ((pb, "prebuffer"), (qb, "postbuffer")):
                    if key in mc and mc[key]:
                        _buffer_spec(buf, mc[key][0])
        pre.append(pb); post.append(qb); m_outages.append(outs); setup_maps.append(setup)
    # transports (mapper.py:1067-1106, :1150-1163)
    t_outages: list = []
    if _has_key(("instance_config", "logistics", "type"), spec_dict):
        lg = ic["logistics"]
        if "amount" not in lg:
            raise InvalidTransportConfig("Transport configuration must include 'amount' key")
        if str(lg.get("type", "")).lower() != "agv":
            raise InvalidTransportConfig(f"Unknown transport type: {lg.get('type')}")
        if _has_key(("instance_config", "outages"), spec_dict):
            for o in ic["outages"]:
                if o["component"] in ["t", "transport", "Transport", "TRANSPORT"]:
                    t_outages.append(_outage(o, sampler))
        T = int(lg["amount"])
    else:
        T = J
    agv_buf_ids = [counter.new_buffer_id() for _ in range(T)]
    if not sbufs:
        sbufs = [{"id": counter.new_buffer_id(), "type": L.BUF_FLEX, "cap": L.CAP_INF, "role": L.ROLE_INPUT},
                 {"id": counter.new_buffer_id(), "type": L.BUF_FLEX, "cap": L.CAP_INF, "role": L.ROLE_OUTPUT}]
    S = len(sbufs)
    # specification (mapper.py:697-793)
    job_op_start, op_machine, op_dur, op_tool_names = [0], [], [], []
    for j, line in enumerate(spec_lines):
        params = [(int(a), int(b)) for a, b in re.findall(r"\((\d+),(\d+)\)", line.split("|")[1])]
        if _has_key(("instance_config", "instance", "tool_usage"), spec_dict):
            tu = next((i for i in ic["instance"]["tool_usage"] if i["job"] == f"j{j}"), None)
            if tu is None:
                raise InvalidToolUsageError(f"j{j}")
            tools_per_op = tu["operation_tools"]
        else:
            tools_per_op = ["tl-0"] * len(params)
        for k, (mach, d) in enumerate(params):
            op_machine.append(mach)
            op_dur.append(_get_time(d, "static", sampler))  # instance.time_behavior is ignored (mapper.py:780-787)
            op_tool_names.append(tools_per_op[k])
        job_op_start.append(len(op_machine))
    if any(m >= M for m in op_machine):
        raise InvalidValue("operation machine", max(op_machine))
    # tools the state machine can ever mount: "tl-0" (initial, mapper.py:396) and the operations' tools.
    # The default setup matrix is all-zero over tl-0..tl-(M-1) (mapper.py:155-171).
    tools = sorted({"tl-0"} | set(op_tool_names))
    nt = len(tools)
    tix = {t: i for i, t in enumerate(tools)}
    default_tools = {f"tl-{i}" for i in range(M)}
    setup_objs: list = []
    for m in range(M):
        for a in tools:
            for b in tools:
                if setup_maps[m] is None:
                    setup_objs.append(_Time(0) if (a in default_tools and b in default_tools) else None)
                else:
                    setup_objs.append(setup_maps[m].get((a, b)))
    # logistics (mapper.py:1178-1199, :570-659)
    place_ids = [f"m-{m}" for m in range(M)] + [b["id"] for b in sbufs]
    pix = {p: i for i, p in enumerate(place_ids)}
    NP = len(place_ids)
    travel_objs: list = [None] * (NP * NP)
    if _has_key(("instance_config", "logistics"), spec_dict):
        lg = ic["logistics"]
        if not _has_key(("instance_config", "logistics", "specification"), spec_dict):
            raise MissingSpecificationError("Logistics specification")
        tb = lg["time_behavior"] if _has_key(("instance_config", "logistics", "time_behavior"), spec_dict) else "static"
        in_id, out_id = sbufs[0]["id"], sbufs[1]["id"]

        def name(n):
            if n.startswith("m-") or n.startswith("b-"):
                return n
            if n.lower() in _INPUT_NAMES:
                return in_id
            if n.lower() in _OUTPUT_NAMES:
                return out_id
            raise UnknownLocationNameError(n)

        for row, col, v in _parse_matrix(lg["specification"]):
            t = _get_time(v, tb, sampler)
            r, c = name(row), name(col)
            if r in pix and c in pix:
                travel_objs[pix[r] * NP + pix[c]] = t
    else:
        travel_objs = [_Time(0) for _ in range(NP * NP)]
    # init state (mapper.py:1249-1407)
    init = spec_dict.get("init_state", {}) or {}
    buf_idx = {}
    for m in range(M):
        buf_idx[pre[m]["id"]] = 3 * m
        buf_idx[post[m]["id"]] = 3 * m + 1
        buf_idx[mbuf_ids[m]] = 3 * m + 2
    for t in range(T):
        buf_idx[agv_buf_ids[t]] = 3 * M + t
    for s, b in enumerate(sbufs):
        buf_idx[b["id"]] = 3 * M + T + s
    location = sbufs[0]["id"]
    job_init = []
    for j in range(J):
        jd = init.get(f"j-{j}", {}) or {}
        for key in jd.keys():
            if key == "operations":
                raise NotImplementedError_()
            if key == "location":
                location = jd["location"]  # sticks for the following jobs too (mapper.py:1252-1267)
        if location not in buf_idx:
            raise InvalidValue("job location", location)
        job_init.append(buf_idx[location])
    if _has_key(("init_state", "components", "machines"), init) or _has_key(("init_state", "components", "buffer"), init):
        raise NotImplementedError_()
    agv_place = []
    for t in range(T):
        loc = f"m-{t % M}"
        ts = init.get(f"t-{t}", {}) or {}
        for key in ts.keys():
            if key == "location":
                loc = ts["location"]
            else:
                raise NotImplementedError_(f"init_state.t-{t}.{key} is not supported by the batched engine")
        if loc not in pix:
            raise ComponentAssociationError(f"t-{t}", "Transport")
        agv_place.append(pix[loc])
    start_time = int(spec_dict["init_state"]["start_time"]) if _has_key(("init_state", "start_time"), spec_dict) else 0

    def arrays(objs):
        time = np.asarray([L.NO_TIME if o is None else o.time for o in objs], dtype=np.int32)
        kind = np.asarray([0 if o is None else o.kind for o in objs], dtype=np.int32)
        if not kind.any():
            return time, L.TimeSpec()
        base = np.asarray([0 if o is None else o.base for o in objs], dtype=np.int32)
        param = np.asarray([0.0 if o is None else o.param for o in objs], dtype=np.float32)
        return time, L.TimeSpec(kind, base, param)

    op_duration, op_spec = arrays(op_dur)
    setup_time, setup_spec = arrays(setup_objs)
    travel_time, travel_spec = arrays(travel_objs)
    m_out_start = np.concatenate([[0], np.cumsum([len(o) for o in m_outages])]).astype(np.int32)
    flat_m = [o for outs in m_outages for o in outs]
    m_ofreq, m_ofreq_spec = arrays([o[0] for o in flat_m])
    m_odur, m_odur_spec = arrays([o[1] for o in flat_m])
    t_ofreq, t_ofreq_spec = arrays([o[0] for o in t_outages])
    t_odur, t_odur_spec = arrays([o[1] for o in t_outages])
    buf_ref = np.zeros(3 * M + T + S, dtype=np.int32)
    for bid, bi in buf_idx.items():
        buf_ref[bi] = int(bid.split("-")[1])
    return L.LoweredInstance(
        description=str(ic.get("description", "")),
        n_jobs=J, n_machines=M, n_transports=T, n_sbuf=S, n_tools=nt,
        job_op_start=job_op_start, op_machine=op_machine, op_tool=[tix[t] for t in op_tool_names],
        op_duration=op_duration,
        pre_type=[b["type"] for b in pre], pre_cap=[_cap(b["cap"]) for b in pre],
        post_type=[b["type"] for b in post], post_cap=[_cap(b["cap"]) for b in post],
        setup_time=setup_time, default_tool=tix["tl-0"],
        sbuf_type=[b["type"] for b in sbufs], sbuf_cap=[_cap(b["cap"]) for b in sbufs],
        sbuf_role=[b["role"] for b in sbufs], buf_ref_id=buf_ref, travel_time=travel_time,
        agv_init_place=agv_place, job_init_buf=job_init, start_time=start_time,
        m_outage_start=m_out_start, m_outage_freq=m_ofreq, m_outage_dur=m_odur,
        t_outage_freq=t_ofreq, t_outage_dur=t_odur,
        lower_bound=_taillard_lower_bound(np.asarray(job_op_start), op_machine, op_duration),
        max_allowed_time=int(np.sum(op_duration)),
        tools=tools, op_spec=op_spec, setup_spec=setup_spec, travel_spec=travel_spec,
        m_ofreq_spec=m_ofreq_spec, m_odur_spec=m_odur_spec, t_ofreq_spec=t_ofreq_spec, t_odur_spec=t_odur_spec,
    )


def _outage(o: dict, sampler):
    """mapper.py:989-1008: (frequency, duration) time objects of one outage spec."""
    t = o["type"]
    if t not in ("maintenance", "repair", "breakdown", "fail", "recharge", "recharging"):
        raise InvalidOutageTypeError(t)
    return (_get_time(None, o["frequency"], sampler), _get_time(None, o["duration"], sampler))


# --------------------------------------------------------------------------- manipulators
def randomize_instance(li: L.LoweredInstance) -> L.LoweredInstance:
    """compiler/manipulators.py:99-150 InstanceRandomizer.manipulate on the lowered tables: per job
    ``np.random.permutation`` of its operations, every duration ``random.randrange(min, max)`` over the
    instance's current durations.  Same calls on the same global generators in the same order as the
    reference, so equal ``np.random.seed`` / ``random.seed`` give the reference's instance.  The derived
    figures (lower bound, max_allowed_time: utils.py:310-352) follow the new tables.  The batched, on-device
    form is ``engine.Batch.randomize``."""
    import dataclasses
    import random

    d = np.asarray(li.op_duration)
    lo, hi = int(d.min()), int(d.max())
    om, ot, od = li.op_machine.copy(), li.op_tool.copy(), li.op_duration.copy()
    for j in range(li.n_jobs):
        a, b = int(li.job_op_start[j]), int(li.job_op_start[j + 1])
        perm = np.random.permutation(b - a)
        om[a:b], ot[a:b] = li.op_machine[a:b][perm], li.op_tool[a:b][perm]
        for k in range(a, b):
            od[k] = random.randrange(lo, hi)  # ValueError for an empty range, like the reference
    return dataclasses.replace(li, op_machine=om, op_tool=ot, op_duration=od, op_spec=L.TimeSpec(),
                               lower_bound=_taillard_lower_bound(np.asarray(li.job_op_start), om, od),
                               max_allowed_time=int(np.sum(od)))


MANIPULATORS = {"DummyManipulator": lambda li: li, "InstanceRandomizer": randomize_instance}


# --------------------------------------------------------------------------- Compiler facade
class Compiler:
    """Same call shape as jobshoplab.compiler.Compiler (compiler.py:16-106); `compile()` returns a
    LoweredInstance (the reference returns ``(InstanceConfig, State)``; both are carried by it)."""

    def __init__(self, config=None, loglevel: Any = "warning", repo: Repository | None = None):
        self.config = config
        self.loglevel = loglevel
        if repo is None:
            if config is None:
                raise InvalidValue("Compiler needs a config or a repo", None)
            name = config.compiler.repo
            if name == "SpecRepository":
                repo = SpecRepository(config.compiler.spec_repository.dir)
            elif name == "DslRepository":
                repo = DslRepository(config.compiler.dsl_repository.dir)
            else:
                raise InvalidValue("compiler.repo", name)
        self.repo = repo
        names = [m for m in (config.compiler.manipulators if config is not None else []) if m is not None]
        for m in names:
            if m not in MANIPULATORS:
                raise InvalidValue("compiler.manipulators", m)  # the reference: AttributeError from getattr
        self.manipulators = tuple(MANIPULATORS[m] for m in names)

    def compile(self, manipulate: bool = True) -> L.LoweredInstance:
        """manipulate=False: the instance as loaded, before compiler.manipulators run (what the reference's
        InstanceRandomizer is handed on every reset: it takes its duration range from THAT instance)."""
        li = lower_dict(self.repo.load_as_dict())
        for fn in (self.manipulators if manipulate else ()):  # compiler.py:98-99
            li = fn(li)
        return li


def compile_spec_text(text: str) -> L.LoweredInstance:
    return lower_dict(SpecRepository.text_to_dict(text))


def compile_dsl_text(text: str) -> L.LoweredInstance:
    return lower_dict(yaml.safe_load(text))
