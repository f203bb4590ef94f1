"""Seam-2 plug-in: the CUDA engine behind the REFERENCE's own middleware slot.

    from jobshoplab import JobShopLabEnv                       # the unmodified reference
    from jobshoplab_b200.middleware import CudaEventBasedBinaryActionMiddleware
    env = JobShopLabEnv(config=cfg, middleware=CudaEventBasedBinaryActionMiddleware)

``CudaEventBasedBinaryActionMiddleware`` honours the contract of
jobshoplab/state_machine/middleware/middleware.py:127-208 (``Middleware``) with the behaviour of
``EventBasedBinaryActionMiddleware`` (:211-443): the reference env builds it through
``DependencyBuilder.state_simulator`` (env/env.py:275-290) with ``loglevel, config, instance, action_factory,
observation_factory`` plus the ``middleware.event_based_binary_action_middleware`` config group, calls
``reset(init_state)`` / ``step(state, action)`` / ``is_truncated()`` and reads ``action_space``.  Every state-machine
step runs in the CUDA engine (a 1-env batch of the logging build); what comes back is a genuine reference
``StateMachineResult`` -- ``State`` dataclasses rebuilt from the engine's canonical digest and per-op log -- so the
reference env's own ``is_done``, reward factory, ``history``, ``info`` and dispatch-rule helpers keep working on it.

The reference must be importable (its objects are what this adapter consumes and produces); nothing else in the
package needs it.  Observation: for the stock factories the engine's packed record is unpacked into the reference's
dict; any other ``observation_factory`` object is simply called on the rebuilt result (user-defined observations,
SURVEY 8(f)4).
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np

from . import abi
from .engine import Batch
from .ref_lowering import Indexer, lower_reference

_STOCK_OBS = {"BinaryActionObservationFactory": abi.OBS_BINARY_ACTION,
              "BinaryOperationArrayObservation": abi.OBS_OPERATION_ARRAY,
              "TasselJsspObservation": abi.OBS_TASSEL}


def host_observation(batch: Batch, idx: int = 0, raw=None) -> dict:
    """Packed observation record of env `idx` -> the dict the reference's factory returns (tuples of bool / int for
    the default factory, observations.py:419-479; float arrays for the others).  `raw`: the env's observation bytes
    when the caller already fetched them (Batch.fetch_host)."""
    if raw is None:
        raw = batch.fetch_host(idx)[1]
    host = {k: v.copy() for k, v in batch.unpack_obs_host(raw).items()}
    if batch.cfg.obs_kind == abi.OBS_BINARY_ACTION:
        for k in ("job_running", "machine_running", "available_jobs"):
            host[k] = tuple(host[k].astype(bool).tolist())
        host["job_executed_on_machine"] = tuple(map(tuple, host["job_executed_on_machine"].astype(bool).tolist()))
        host["job_progression"] = tuple(host["job_progression"].tolist())
        host["machine_progression"] = tuple(host["machine_progression"].tolist())
    return host


class _Ref:
    """The reference's types, imported on first use."""

    def __init__(self):
        from jobshoplab.state_machine import time_machines
        from jobshoplab.types import action_types, state_types
        from jobshoplab.utils import exceptions

        self.st, self.at, self.tm, self.ex = state_types, action_types, time_machines, exceptions
        s = state_types
        self.m_state = (s.MachineStateState.IDLE, s.MachineStateState.SETUP, s.MachineStateState.WORKING,
                        s.MachineStateState.OUTAGE)
        self.t_state = (s.TransportStateState.IDLE, s.TransportStateState.WORKING, s.TransportStateState.PICKUP,
                        s.TransportStateState.TRANSIT, s.TransportStateState.OUTAGE, s.TransportStateState.WAITINGPICKUP)
        self.o_state = (s.OperationStateState.IDLE, s.OperationStateState.PROCESSING, s.OperationStateState.DONE)


class StateBuilder:
    """Canonical digest (layout: DESIGN.md) -> reference ``State`` + ``possible_transitions``."""

    def __init__(self, instance, init_state, ix: Indexer, ref: _Ref):
        self.instance, self.ix, self.ref = instance, ix, ref
        self.init_state = init_state
        caps = {}
        for m in instance.machines:
            for b in (m.prebuffer, m.postbuffer, m.buffer):
                caps[b.id] = b.capacity
        for t in instance.transports:
            caps[t.buffer.id] = t.buffer.capacity
        for b in instance.buffers:
            caps[b.id] = b.capacity
        self.caps = caps
        self.init_store = {}
        for m in init_state.machines:
            for b in (m.prebuffer, m.postbuffer, m.buffer):
                self.init_store[b.id] = tuple(b.store)
        for t in init_state.transports:
            self.init_store[t.buffer.id] = tuple(t.buffer.store)
        for b in init_state.buffers:
            self.init_store[b.id] = tuple(b.store)
        self.ops = [[(o.id, o.machine) for o in job.operations] for job in instance.instance.specification]

    def _time(self, v):
        return self.ref.st.Time(int(v)) if v >= 0 else self.ref.st.NoTime()

    def _buffer(self, bid, store):
        s = self.ref.st
        store = tuple(store)
        if not store:
            state = s.BufferStateState.EMPTY
        elif len(store) == self.caps[bid] and store != self.init_store.get(bid):
            state = s.BufferStateState.FULL  # buffer_type_utils.py:96-97: the last put filled it
        else:
            state = s.BufferStateState.NOT_EMPTY
        return s.BufferState(id=bid, state=state, store=store)

    def build(self, digest: np.ndarray, prev_state=None):
        s, ix = self.ref.st, self.ix
        d = [int(x) for x in digest]
        p = 0

        def take(n=1):
            nonlocal p
            v = d[p:p + n]
            p += n
            return v if n != 1 else v[0]

        def store():
            n = take()
            return [ix.jobs[x] for x in (take(n) if n != 1 else [take()])] if n else []

        time, J, M, T, S = take(5)
        jobs = []
        for j in range(J):
            _nd, _run, loc = take(3)
            ops = []
            for oid, mid in self.ops[j]:
                st, a, b = take(3)
                ops.append(s.OperationState(id=oid, start_time=self._time(a), end_time=self._time(b), machine_id=mid,
                                            operation_state_state=self.ref.o_state[st]))
            jobs.append(s.JobState(id=ix.jobs[j], operations=tuple(ops), location=ix.buffer_ids[loc]))
        machines = []
        for m, (cfg, ini) in enumerate(zip(self.instance.machines, self.init_state.machines)):
            st, occ, tool = take(3)
            pre, buf, post = store(), store(), store()
            n_out = len(cfg.outages)
            machines.append(dict(cfg=cfg, ini=ini, st=st, occ=occ, tool=tool, pre=pre, buf=buf, post=post, n_out=n_out))
        transports = []
        for t, (cfg, ini) in enumerate(zip(self.instance.transports, self.init_state.transports)):
            transports.append(dict(cfg=cfg, ini=ini, v=take(12)))
        sbufs = [store() for _ in range(S)]

        def outages(ini_outages):
            out = []
            for o in ini_outages:
                active, a, b = take(3)
                if active:
                    out.append(s.OutageState(id=o.id, active=s.OutageActive(start_time=self._time(a), end_time=self._time(b))))
                else:
                    out.append(s.OutageState(id=o.id, active=s.OutageInactive(last_time_active=self._time(a))))
            return tuple(out)

        m_states = []
        for m in machines:
            cfg, ini = m["cfg"], m["ini"]
            m_states.append(s.MachineState(
                id=cfg.id, buffer=self._buffer(cfg.buffer.id, m["buf"]), occupied_till=self._time(m["occ"]),
                prebuffer=self._buffer(cfg.prebuffer.id, m["pre"]), postbuffer=self._buffer(cfg.postbuffer.id, m["post"]),
                state=self.ref.m_state[m["st"]], mounted_tool=ix.tools[m["tool"]], outages=outages(ini.outages),
                resources=ini.resources))
        t_states = []
        prev_t = {t.id: t for t in prev_state.transports} if prev_state is not None else {}
        for tr in transports:
            cfg, ini = tr["cfg"], tr["ini"]
            st, kind, occ, td_buf, td_comp, td_job, carried, tjob, lk, l0, l1, l2 = tr["v"]
            if kind == 2:
                occ_t = s.TimeDependency(
                    buffer_id=ix.buffer_ids[td_buf], job_id=ix.jobs[occ],
                    transition=self.ref.at.ComponentTransition(
                        component_id=ix.transports[td_comp], new_state=s.TransportStateState.WAITINGPICKUP,
                        job_id=ix.jobs[td_job] if td_job >= 0 else None))
            else:
                occ_t = self._time(occ if kind == 1 else -1)
            state = self.ref.t_state[st]
            if lk == 0:
                # a place: 1.0 until the AGV first moves (mapper.py:420), 0 after a drop-off (manipulate.py:100)
                before = prev_t.get(cfg.id)
                if before is not None and isinstance(before.location.location, str) and before.location.location == ix.place_id(l0):
                    progress = before.location.progress
                else:
                    progress = ini.location.progress if prev_state is None else 0
                loc = s.TransportLocation(progress=progress, location=ix.place_id(l0))
            else:
                # a route: 0.0 from the assignment (core_utils.py:213), 0.5 once in transit (handler.py:791)
                progress = 0.5 if state == s.TransportStateState.TRANSIT else 0.0
                loc = s.TransportLocation(progress=progress, location=(ix.place_id(l0), ix.buffer_ids[l1], ix.place_id(l2)))
            t_states.append(s.TransportState(
                state=state, id=cfg.id, occupied_till=occ_t,
                buffer=self._buffer(cfg.buffer.id, [ix.jobs[carried]] if carried >= 0 else []), location=loc,
                outages=None, transport_job=ix.jobs[tjob] if tjob >= 0 else None))
        for i, (tr, ts) in enumerate(zip(transports, t_states)):
            t_states[i] = replace(ts, outages=outages(tr["ini"].outages))
        buffers = tuple(self._buffer(b.id, st) for b, st in zip(self.instance.buffers, sbufs))
        n_pos = take()
        poss = []
        for _ in range(n_pos):
            k, comp, ns, job = take(4)
            poss.append(self.ref.at.ComponentTransition(
                component_id=ix.machines[comp] if k == 0 else ix.transports[comp],
                new_state=self.ref.m_state[ns] if k == 0 else self.ref.t_state[ns],
                job_id=ix.jobs[job] if job >= 0 else None))
        state = s.State(jobs=tuple(jobs), time=s.Time(time), machines=tuple(m_states), transports=tuple(t_states),
                        buffers=buffers)
        return state, tuple(poss)


class CudaEventBasedBinaryActionMiddleware:
    """Drop-in for ``EventBasedBinaryActionMiddleware`` (middleware.py:211-443) running on the CUDA engine."""

    # The reference env builds a NEW middleware object on every env.reset (env.py:518-524); device batches are
    # therefore kept per lowered instance at class level, so that a reset of an unchanged instance allocates nothing.
    _batches: dict = {}
    _MAX_BATCHES = 4

    def __init__(self, loglevel, config, instance, action_factory, observation_factory, truncation_joker: int = 5,
                 truncation_active: bool = False, state_machine_step=None, device: int = 0, *args, **kwargs):
        self.loglevel, self.config, self.instance = loglevel, config, instance
        self.action_factory, self.observation_factory = action_factory, observation_factory
        self.action_space = action_factory.action_space
        self.truncation_joker, self._init_truncation_joker = truncation_joker, truncation_joker
        self.truncation_active = truncation_active
        self.device = device
        self._ref = _Ref()
        self._ix = Indexer(instance)
        self._batch = None
        self._lowered_key = None
        self._truncated = False
        kind = _STOCK_OBS.get(type(observation_factory).__name__)
        self._obs_kind = abi.OBS_BINARY_ACTION if kind is None else kind
        self._engine_obs = kind is not None

    # ------------------------------------------------------------------ helpers
    def _cfg(self) -> dict:
        rw = self.config.reward_factory.binary_action_jssp_reward
        return dict(allow_early_transport=self.config.state_machine.allow_early_transport,
                    truncation_joker=self._init_truncation_joker, truncation_active=self.truncation_active,
                    obs_kind=self._obs_kind, sparse_bias=rw.sparse_bias, dense_bias=rw.dense_bias,
                    truncation_bias=rw.truncation_bias)

    def _raise_for(self, status: int):
        ex = self._ref.ex
        if status == abi.ST_BUFFER_FULL:
            raise ex.BufferFullError("engine status %d" % status)
        if status == abi.ST_UNSUCCESSFUL:
            raise ex.UnsuccessfulStateMachineResult()
        if status == abi.ST_NO_POSSIBLE:
            raise ex.InvalidValue("state.possible_transitions", (), "No possible transitions")
        if status == abi.ST_ENV_DONE:
            raise ex.EnvDone()
        if status >= abi.ST_ACTION_OUT_OF_SPACE:
            raise ex.InvalidValue("engine status", status, "the state machine raised")

    def _observation(self, result, done: bool):
        if self._engine_obs:
            return host_observation(self._batch, 0, self._obs_raw)
        return self.observation_factory.make(result, done)

    def _fetch(self):
        """One device->host copy per call: (status, terminated, truncated) of the record; keeps the observation bytes."""
        rec, self._obs_raw = self._batch.fetch_host(0)
        return int(rec[26]), bool(rec[24]), bool(rec[25])

    def _result(self, action, success=True, message="Success", prev=None):
        state, poss = self._builder.build(self._batch.get_state(0), prev_state=prev.state if prev is not None else None)
        return self._ref.st.StateMachineResult(state=state, sub_states=(), action=action, success=success,
                                               message=message, possible_transitions=poss)

    # ------------------------------------------------------------------ Middleware contract
    def reset(self, init_state):
        """middleware.py:264-290: one dummy step with the empty action, first observation, joker reset."""
        li = lower_reference(self.instance, init_state, self._ix)
        key = (li.op_machine.tobytes(), li.op_duration.tobytes(), li.travel_time.tobytes(), li.setup_time.tobytes(),
               li.start_seeds.tobytes(), tuple(li.job_init_buf), tuple(li.agv_init_place))
        key = (key, tuple(sorted(self._cfg().items())), self.device)
        cache = type(self)._batches
        cached = cache.get(key)
        owner = getattr(cached, "_owner", lambda: None)() if cached is not None else None
        if cached is not None and cached.h is not None and owner in (None, self):
            self._batch = cached  # same tables + config and its previous middleware is gone: reuse the device memory
        else:
            if cached is None or cached.h is None:
                while len(cache) >= self._MAX_BATCHES:
                    cache.pop(next(iter(cache)))
            self._batch = Batch(li, 1, device=self.device, record_ops=True, **self._cfg())
            if cached is None or cached.h is None:
                cache[key] = self._batch
        import weakref
        self._batch._owner = weakref.ref(self)
        self._lowered_key = key
        if not li.is_deterministic:  # the reference's own numpy start seeds: same sample streams
            import torch
            self._seeds = torch.from_numpy(np.ascontiguousarray(li.start_seeds.reshape(1, -1))).to(f"cuda:{self.device}")
            self._batch.set_start_seeds(self._seeds)
        self._builder = StateBuilder(self.instance, init_state, self._ix, self._ref)
        import torch
        self._acts = torch.tensor([[0], [1]], dtype=torch.uint8, device=f"cuda:{self.device}").unbind(0)  # resident action bytes
        self._batch.reset()
        status, terminated, _ = self._fetch()
        self._raise_for(status)
        self.truncation_joker = self._init_truncation_joker
        self._truncated = False
        result = self._result(self.action_factory.get_dummy_action(), message="Success")
        if not result.possible_transitions and terminated:
            result = replace(result, message="Done")
        return result, self._observation(result, False)

    def step(self, state, action):
        """middleware.py:384-431 (+ :304-368 no-op handling) with the state machine on the GPU."""
        ref = self._ref
        if not isinstance(state, ref.st.StateMachineResult):
            raise ref.ex.InvalidValue("State must be of type StateMachineResult")
        act = self.action_factory.interpret(action, state)  # raises InvalidValue / ActionOutOfActionSpace like the stock one
        no_op = act.action_factory_info == ref.at.ActionFactoryInfo.NoOperation
        self._batch.step(self._acts[0 if no_op else 1])
        status, terminated, truncated = self._fetch()
        self._raise_for(status)
        self._truncated = truncated and status != abi.ST_STEP_FAILED
        if status == abi.ST_STEP_FAILED:  # validation error: the old state, success=False (state.py:202-210)
            result = ref.st.StateMachineResult(state=state.state, sub_states=state.sub_states, action=act, success=False,
                                               message="Transition errors", possible_transitions=())
            return result, self._observation(result, True)
        if no_op and len(state.possible_transitions) > 1:
            # :355-368 -- nothing ran, the first candidate is dropped
            result = ref.st.StateMachineResult(
                state=state.state, sub_states=state.sub_states,
                action=ref.at.Action(tuple(), action_factory_info=ref.at.ActionFactoryInfo.NoOperation,
                                     time_machine=ref.tm.jump_to_event),
                success=True, message="NO OPERATION", possible_transitions=state.possible_transitions[1:])
        else:
            if no_op:
                act = replace(act, time_machine=ref.tm.force_jump_to_event)  # :336
            result = self._result(act, prev=state)
            if not result.possible_transitions and terminated:
                result = replace(result, message="Done")
        if no_op and len(state.possible_transitions) == 1 and result.possible_transitions:
            self.truncation_joker = self._init_truncation_joker  # informational; the engine holds the real counter
        return result, self._observation(result, len(result.possible_transitions) == 0)

    def is_truncated(self) -> bool:
        """middleware.py:433-443: truncation joker below zero (the engine counts it)."""
        return self._truncated

    def close(self):
        """Frees the device batch of this middleware's instance (and drops it from the class-level cache)."""
        if self._batch is not None:
            for k, b in list(type(self)._batches.items()):
                if b is self._batch:
                    del type(self)._batches[k]
            self._batch.close()
            self._batch = None
