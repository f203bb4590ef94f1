"""Read-only views of an env's state rebuilt from the engine's canonical digest, shaped like the
reference objects user code touches between steps: ``env.state.state.time.time``,
``env.state.possible_transitions[0].component_id / .job_id / .new_state``, job locations, machine and
transport states (types/state_types.py:44-267, action_types.py:11-34).  Host-side convenience for
dispatch-rule code, render() and debugging -- never on the step path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

M_STATES = ("Idle", "Setup", "Working", "Outage")
T_STATES = ("Idle", "Working", "Pickup", "Transit", "Outage", "WaitingPickup")
O_STATES = ("Idle", "Processing", "Done")


@dataclass(frozen=True)
class Time:
    time: int


@dataclass(frozen=True)
class ComponentTransition:
    component_id: str
    new_state: str
    job_id: Optional[str]


@dataclass(frozen=True)
class OperationView:
    id: str
    machine_id: str
    state: str
    start_time: Optional[int]
    end_time: Optional[int]


@dataclass(frozen=True)
class JobView:
    id: str
    location: str
    operations: tuple


@dataclass(frozen=True)
class MachineView:
    id: str
    state: str
    occupied_till: Optional[int]
    mounted_tool: str
    prebuffer: tuple
    buffer: tuple
    postbuffer: tuple


@dataclass(frozen=True)
class TransportView:
    id: str
    state: str
    occupied_till: Optional[int]
    transport_job: Optional[str]
    carried: Optional[str]
    location: object


@dataclass(frozen=True)
class StateView:
    time: Time
    jobs: tuple
    machines: tuple
    transports: tuple
    buffers: tuple


@dataclass(frozen=True)
class ResultView:
    state: StateView
    possible_transitions: tuple


def parse_digest(li, d: np.ndarray) -> ResultView:
    """Canonical digest (layout in DESIGN.md) -> ResultView, with the reference's string ids."""
    d = [int(x) for x in d]
    p = 0

    def take(n=1):
        nonlocal p
        v = d[p:p + n]
        p += n
        return v if n > 1 else v[0]

    M, T = li.n_machines, li.n_transports

    def buf_name(b):
        return f"b-{int(li.buf_ref_id[b])}"

    def place_name(x):
        return f"m-{x}" if x < M else buf_name(3 * M + T + (x - M))

    def opt(t):
        return None if t < 0 else t

    time, J, M_, T_, S = take(5)
    jobs = []
    for j in range(J):
        _n_done, _run, loc = take(3)
        ops = []
        for k in range(int(li.job_op_start[j + 1] - li.job_op_start[j])):
            st, a, b = take(3)
            o = int(li.job_op_start[j]) + k
            ops.append(OperationView(f"o-{j}-{k}", f"m-{int(li.op_machine[o])}", O_STATES[st], opt(a), opt(b)))
        jobs.append(JobView(f"j-{j}", buf_name(loc), tuple(ops)))

    def store():
        n = take()
        return tuple(f"j-{x}" for x in (take(n) if n > 1 else ([take()] if n == 1 else [])))

    machines = []
    for m in range(M_):
        st, occ, tool = take(3)
        pre, buf, post = store(), store(), store()
        machines.append(MachineView(f"m-{m}", M_STATES[st], opt(occ), li.tools[tool] if li.tools else f"tl-{tool}",
                                    pre, buf, post))
    transports = []
    for t in range(T_):
        st, kind, occ, _tb, _tc, _tj, carried, tjob, lk, l0, l1, l2 = take(12)
        loc = place_name(l0) if lk == 0 else (place_name(l0), buf_name(l1), place_name(l2))
        transports.append(TransportView(f"t-{t}", T_STATES[st], occ if kind == 1 else None,
                                        None if tjob < 0 else f"j-{tjob}", None if carried < 0 else f"j-{carried}", loc))
    buffers = tuple((buf_name(3 * M + T + s), store()) for s in range(S))
    n_out = len(li.m_outage_freq) + T_ * len(li.t_outage_freq)
    take(3 * n_out) if n_out else None
    n_pos = take()
    poss = []
    for _ in range(n_pos):
        kind, comp, ns, job = take(4)
        poss.append(ComponentTransition(f"m-{comp}" if kind == 0 else f"t-{comp}",
                                        M_STATES[ns] if kind == 0 else T_STATES[ns], None if job < 0 else f"j-{job}"))
    return ResultView(StateView(Time(time), tuple(jobs), tuple(machines), tuple(transports), buffers), tuple(poss))


def gantt_data(li, view: ResultView, legs: np.ndarray) -> dict:
    """rendering/gant_dashboard.py:414-429 map_states_to_gant_data: what the reference's dashboard draws.
    `view`: the current state (parse_digest), `legs`: the transport-task log (jsl_get_leg_log, int32 [n][5]).

    schedules  (:250-296) DONE operations of every job, then its PROCESSING one, from the last state;
    transports (:299-359) per AGV in id order, one bar per run of recorded non-idle states sharing a route
               (consecutive tasks with the same route therefore merge: start of the first, job / end of the last);
    buffer     (:362-376) not implemented upstream: ()."""
    M, T = li.n_machines, li.n_transports

    def buf_name(b):
        return f"b-{int(li.buf_ref_id[b])}"

    def place_name(x):
        return f"m-{x}" if x < M else buf_name(3 * M + T + (x - M))

    schedules = []
    for job in view.state.jobs:
        for op in job.operations:
            if op.state == "Done":
                schedules.append({"job": job.id, "start": op.start_time, "end": op.end_time, "id": op.machine_id,
                                  "type": "Schedule", "meta_info": None})
        active = next((op for op in job.operations if op.state == "Processing"), None)
        if active is not None:
            schedules.append({"job": job.id, "start": active.start_time, "end": active.end_time, "id": active.machine_id,
                              "type": "Schedule", "meta_info": None})
    transports = []
    legs = np.asarray(legs, dtype=np.int64).reshape(-1, 5)
    for t in range(T):
        prev = None
        for agv, job, start, end, route in legs[legs[:, 0] == t]:
            loc = (place_name(int(route & 0xff)), buf_name(int((route >> 8) & 0xffff)), place_name(int((route >> 24) & 0xff)))
            if prev is not None and prev["_loc"] == loc:
                prev["job"], prev["end"] = f"j-{job}", (None if end < 0 else int(end))
                continue
            prev = {"type": "Transport", "id": f"t-{t}", "job": f"j-{job}", "start": int(start),
                    "end": None if end < 0 else int(end), "meta_info": f"route: {loc}", "_loc": loc}
            transports.append(prev)
    for d in transports:
        del d["_loc"]
    return {"schedules": tuple(schedules), "transports": tuple(transports), "buffer": ()}
