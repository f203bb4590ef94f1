"""JobShopLabEnv / BatchedJobShopLabEnv: the reference's env surface over the CUDA engine.

``JobShopLabEnv`` keeps the call shape of jobshoplab/env/env.py:293-618 -- ``JobShopLabEnv(config, seed,
compiler, ...)``, ``reset(seed) -> (obs, info)``, ``step(action) -> (obs, reward, terminated, truncated,
info)``, attributes ``instance, state, done, terminated, truncated, lower_bound, max_allowed_time,
num_operations, observation_space, action_space`` -- and runs as a 1-env batch on the GPU.  The per-env
status code of the engine is turned back into the exception class the reference raises out of
``env.step`` (exceptions.STATUS_TO_EXCEPTION).

``BatchedJobShopLabEnv`` is the vectorised form the reference never shipped (its env_wrapper.py is empty):
B envs, device tensors in and out, optional auto-reset, fused policy rollouts.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import abi, compiler as _compiler, spaces as _spaces
from .config import Config
from .engine import Batch
from .exceptions import (STATUS_TO_EXCEPTION, ConfigurationError, EnvDone, InvalidValue)
from .lowered import LoweredInstance
from .state_view import parse_digest

_OBS_KINDS = {"BinaryActionObservationFactory": abi.OBS_BINARY_ACTION,
              "BinaryOperationArrayObservation": abi.OBS_OPERATION_ARRAY,
              "TasselJsspObservation": abi.OBS_TASSEL,
              # extension: the raw state record, for user-defined observation / reward functions on the torch side
              "StateObservation": abi.OBS_STATE}


class Discrete:
    """Stand-in for gymnasium.spaces.Discrete(2) when gymnasium is not installed."""

    def __init__(self, n, start=0):
        self.n, self.start = int(n), int(start)

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and not isinstance(x, bool) and self.start <= int(x) < self.start + self.n

    def sample(self):
        return int(np.random.randint(self.start, self.start + self.n))


def _cfg_kwargs(config: Config) -> dict:
    """Config groups -> jsl_env_cfg fields (env.py:186-290 resolves the same names to classes)."""
    if config.env.middleware != "EventBasedBinaryActionMiddleware":
        raise ConfigurationError("env.middleware", config.env.middleware)
    if config.env.action_factory != "BinaryJobActionFactory":
        raise ConfigurationError("env.action_factory", config.env.action_factory)
    if config.env.reward_factory != "BinaryActionJsspReward":
        raise ConfigurationError("env.reward_factory", config.env.reward_factory)
    if config.env.observation_factory not in _OBS_KINDS:
        raise ConfigurationError("env.observation_factory", config.env.observation_factory)
    mw = config.middleware.event_based_binary_action_middleware
    rw = config.reward_factory.binary_action_jssp_reward
    return dict(allow_early_transport=config.state_machine.allow_early_transport,
                truncation_joker=mw.truncation_joker, truncation_active=mw.truncation_active,
                obs_kind=_OBS_KINDS[config.env.observation_factory], sparse_bias=rw.sparse_bias,
                dense_bias=rw.dense_bias, truncation_bias=rw.truncation_bias)


def _lower(config: Config, compiler, manipulate: bool = True) -> LoweredInstance:
    if isinstance(compiler, LoweredInstance):
        return compiler
    if compiler is None:
        compiler = _compiler.Compiler(config)
    try:
        out = compiler.compile(manipulate=manipulate) if not manipulate else compiler.compile()
    except TypeError:  # a user compiler without the keyword
        out = compiler.compile()
    if not isinstance(out, LoweredInstance):
        raise InvalidValue("compiler", type(out), "compile() must return a LoweredInstance "
                           "(use jobshoplab_b200.compiler.Compiler)")
    return out


def _same_instance(a: LoweredInstance, b: LoweredInstance) -> bool:
    """Equal lowered tables: the device batch built for `a` serves `b` (reset then costs one kernel, no allocation)."""
    import dataclasses

    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if f.name.endswith("_spec"):
            for k in ("kind", "base", "param"):
                xx, yy = getattr(x, k), getattr(y, k)
                if (xx is None) != (yy is None) or (xx is not None and not np.array_equal(xx, yy)):
                    return False
        elif isinstance(x, np.ndarray):
            if not np.array_equal(x, y):
                return False
        elif f.name != "description" and x != y:
            return False
    return True


class BatchedJobShopLabEnv:
    """B envs of one instance (or of same-shaped instances) on one GPU."""

    def __init__(self, config: Config | None = None, num_envs: int = 1, compiler=None, device: int = 0, seed: int = 0,
                 instances=None, variants=None, record_ops: bool = False, auto_reset: bool = False,
                 observation_fn=None, reward_fn=None):
        """observation_fn / reward_fn: user-defined factories (the reference's ObservationFactory / RewardFactory ABCs,
        observations.py:31-76, rewards.py:11-58, in batched form): callables ``fn(views, out) -> tensor`` over the
        views of the configured observation record (use env.observation_factory = "StateObservation" to get the
        whole state: Batch.unpack_state) and the step scalars ``out``; their results replace the observation /
        reward that reset() / step() return.  (A device-side reward hook: csrc/jsl_engine.cuh "User hook".)"""
        self.config = config or Config()
        self.observation_fn, self.reward_fn = observation_fn, reward_fn
        # compiler.manipulators = [InstanceRandomizer]: the reference recompiles -- and so re-randomises -- the
        # instance on every env.reset (env.py:491).  Batched: every env owns a variant of the operation
        # tables, redrawn on the device whenever that env is reset (engine.Batch.randomize).  The manipulator takes
        # its duration range from the instance it is handed, i.e. the UN-manipulated one (manipulators.py:122-125).
        self.randomize = "InstanceRandomizer" in [m for m in self.config.compiler.manipulators if m]
        lowered = instances if instances is not None else _lower(self.config, compiler, manipulate=not self.randomize)
        if self.randomize and variants is None and instances is None:
            li = lowered
            tile = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.int32), (num_envs, len(a))))
            variants = (tile(li.op_machine), tile(li.op_duration), np.full(num_envs, li.lower_bound, np.int32),
                        np.full(num_envs, li.max_allowed_time, np.int32))
        self.batch = Batch(lowered, num_envs, device=device, seed=seed, variants=variants, record_ops=record_ops,
                           **_cfg_kwargs(self.config))
        if self.randomize and self.batch.n_variants != num_envs:
            # the redraw mask is per env: with shared variants, resetting one env would rewrite tables under others
            raise ConfigurationError("compiler.manipulators", "InstanceRandomizer",
                                     f"needs one variant per env (got {self.batch.n_variants} variants for {num_envs} envs)")
        self.instance = self.batch.lowered[0]
        self.num_envs = num_envs
        self.auto_reset = auto_reset
        self.action_space = _spaces.action_space()
        self.observation_space = _spaces.observation_space(self.batch.cfg.obs_kind, self.instance.n_jobs,
                                                           self.instance.n_machines, self.instance.n_ops)
        self._seed, self._epoch = int(seed), 0
        d = np.asarray(self.instance.op_duration)
        self._dur_range = (int(d.min()), int(d.max()))  # of the instance handed to the manipulator
        if self.randomize:
            self._redraw()  # the reference's constructor compiles (and randomises) once as well

    def _redraw(self, mask=None):
        self._epoch += 1
        self.batch.randomize(seed=(self._seed * 0x9E3779B97F4A7C15 + self._epoch) & 0xFFFFFFFFFFFFFFFF,
                             min_duration=self._dur_range[0], max_duration=self._dur_range[1], mask=mask)

    def reset(self, mask=None):
        if self.randomize:
            self._redraw(mask)
        out = self.batch.reset(mask)
        views = self.batch.unpack_obs(out.obs)
        return (self.observation_fn(views, out) if self.observation_fn else views), out

    def step(self, actions):
        """actions: uint8 CUDA tensor [B].  Returns (obs dict of views, reward, terminated, truncated, out).
        With auto_reset, every env that ended -- terminated, truncated, or dead with an escaped exception
        (status >= JSL_ST_STEP_FAILED) -- is reset at once: reward / terminated / truncated / makespan are the finished
        step's, observation / time / status / candidate the fresh episode's (out.ended: the status the episode ended
        with, when the reset happened inside the kernel)."""
        # auto-reset is fused into the step kernel (JSL_STEP_AUTO_RESET: no host round trip, no second launch) unless
        # the instances are redrawn at every reset, which needs jsl_batch_randomize between the step and the reset
        fused = self.auto_reset and not self.randomize
        out = self.batch.step(actions, auto_reset=fused)
        if self.auto_reset and not fused:
            done = (out.terminated | out.truncated | (out.status >= abi.ST_STEP_FAILED).to(out.terminated.dtype)).contiguous()
            if bool(done.any()):
                final = self.batch.record.clone()
                if self.randomize:
                    self._redraw(done)
                self.batch.reset(done)
                # reward / makespan / flags of the finished step survive the reset; observation, time, status and
                # candidate are the fresh episode's
                rec = self.batch.record
                rec[:, 0:8] = final[:, 0:8]; rec[:, 12:16] = final[:, 12:16]; rec[:, 24:26] = final[:, 24:26]
        views = self.batch.unpack_obs(out.obs)
        reward = self.reward_fn(views, out) if self.reward_fn else out.reward
        return (self.observation_fn(views, out) if self.observation_fn else views), reward, out.terminated, out.truncated, out

    # gymnasium.vector / SB3 VecEnv style split step: the launch is asynchronous on the
    # current CUDA stream anyway, step_wait only hands out the views
    def step_async(self, actions):
        self._pending = self.step(actions)

    def step_wait(self):
        out, self._pending = self._pending, None
        return out

    @property
    def single_action_space(self):
        return self.action_space

    def rollout(self, policy="random", **kw):
        return self.batch.rollout(policy, **kw)

    def state(self, idx: int = 0):
        return parse_digest(self.instance, self.batch.get_state(idx))

    def close(self):
        self.batch.close()


class _ActionView:
    """`.transitions` of the Action a history entry was made with (types/action_types.py:36-46)."""

    __slots__ = ("transitions",)

    def __init__(self, transitions):
        self.transitions = transitions


class HistoryEntry:
    """What user code reads of a ``StateMachineResult`` in ``env.history`` (env.py:404-411, :446): the action's
    transitions (empty = no-op), success, message and the time after the step.  States are not kept per step --
    ``env.state`` rebuilds the current one, ``env.gantt_data()`` / ``render()`` work from the engine's logs."""

    __slots__ = ("action", "success", "message", "time")

    def __init__(self, transitions, success, message, time):
        self.action, self.success, self.message, self.time = _ActionView(transitions), success, message, time


def _factory_name(arg, default):
    if arg is None:
        return default
    if isinstance(arg, str):
        return arg
    return getattr(arg, "__name__", None) or type(arg).__name__


class JobShopLabEnv:
    """Single-env drop-in (env.py:293-618).  `compiler` may be a jobshoplab_b200.compiler.Compiler, a
    LoweredInstance, or None (built from config.compiler like the reference).  The factory / middleware arguments
    take what the reference takes -- a class, an instance or (as in the config) a name -- as long as it names one of
    the stock classes the engine implements; for user-defined factories plug the engine into the reference's own env
    instead (jobshoplab_b200.middleware.CudaEventBasedBinaryActionMiddleware)."""

    metadata = {"render_modes": ["normal", "dashboard", "debug"]}

    def __init__(self, config: Config | None = None, seed=None, compiler=None, observation_factory=None,
                 reward_factory=None, middleware=None, action_factory=None, render_backend=None, loglevel=None,
                 device: int = 0):
        self.config = config or Config()
        stock = {"observation_factory": tuple(_OBS_KINDS), "reward_factory": ("BinaryActionJsspReward",),
                 "middleware": ("EventBasedBinaryActionMiddleware", "CudaEventBasedBinaryActionMiddleware"),
                 "action_factory": ("BinaryJobActionFactory",)}
        for name, arg in (("observation_factory", observation_factory), ("reward_factory", reward_factory),
                          ("middleware", middleware), ("action_factory", action_factory)):
            if arg is None:
                continue
            cls = _factory_name(arg, None)
            if cls not in stock[name]:
                raise ConfigurationError(name, cls, "the CUDA engine implements the reference's stock classes "
                                         f"{stock[name]}; for a user-defined one use the reference env with "
                                         "jobshoplab_b200.middleware.CudaEventBasedBinaryActionMiddleware")
            if name == "observation_factory":
                self.config.env.observation_factory = cls
        self._compiler = compiler
        self._device = device
        self.seed = seed
        self.episode_counter = -1
        self._batch = None
        self.render_backend = render_backend
        self.current_observation = self.reset(seed)

    # ------------------------------------------------------------------ reference attributes
    @property
    def state(self):
        return parse_digest(self.instance, self._batch.get_state(0))

    @property
    def num_operations(self):
        return self.instance.n_ops

    def _fetch(self):
        """One device->host copy per step: the env's 32-byte jsl_step_record and its observation record."""
        from .middleware import host_observation
        rec, obs = self._batch.fetch_host(0)
        raw = rec.tobytes()
        reward = float(np.frombuffer(raw, np.float64, 1, 0)[0])
        time, makespan = (int(x) for x in np.frombuffer(raw, np.int32, 2, 8))
        cand = tuple(int(x) for x in np.frombuffer(raw, np.int16, 4, 16))
        return (reward, time, makespan, cand, bool(raw[24]), bool(raw[25]), raw[26]), host_observation(self._batch, 0, obs)

    def _get_info(self, no_op: bool) -> dict:
        return {"no_op": no_op, "terminated": self.terminated, "truncated": self.truncated,
                "makespan": self._time if self.terminated else None, "no_op_count": self._no_op_count}

    def _cand_transition(self, cand):
        from .state_view import ComponentTransition, M_STATES, T_STATES
        kind, comp, job, n = cand
        if n <= 0:
            return None
        return ComponentTransition(f"m-{comp}" if kind == 0 else f"t-{comp}", M_STATES[1] if kind == 0 else T_STATES[1],
                                   f"j-{job}")

    def reset(self, seed=None):
        self.episode_counter += 1
        if seed is None:
            seed = self.seed
        self.seed = seed + self.episode_counter if seed is not None else None
        instance = _lower(self.config, self._compiler)  # the reference recompiles on every reset (env.py:491)
        if self._batch is None or not _same_instance(instance, self.instance):
            if self._batch is not None:
                self._batch.close()
            self._batch = Batch(instance, 1, device=self._device, seed=self.seed or 0, record_ops=True,
                                **_cfg_kwargs(self.config))
            # the two legal actions as resident device bytes: a step uploads nothing
            self._acts = self._batch.torch.tensor([[0], [1]], dtype=self._batch.torch.uint8, device=f"cuda:{self._device}").unbind(0)
        self.instance = instance
        self.lower_bound = self.instance.lower_bound
        self.max_allowed_time = self.instance.max_allowed_time
        self.action_space = Discrete(2)
        self.observation_space = _spaces.observation_space(self._batch.cfg.obs_kind, instance.n_jobs, instance.n_machines,
                                                           instance.n_ops)
        self._batch.reset()
        (_, self._time, _, self._cand, _, _, status), obs = self._fetch()
        self._raise_for(status)
        self.terminated = self.truncated = self.done = False
        self._no_op_count = 0
        self.history: tuple = tuple()
        return obs, {"env_seed": self.seed}

    def _raise_for(self, status: int):
        exc = STATUS_TO_EXCEPTION.get(status)
        if exc is None:
            return
        if exc is EnvDone:
            raise EnvDone()
        if exc is InvalidValue:
            raise InvalidValue("state.possible_transitions", (), "No possible transitions")
        raise exc(*(["engine status %d" % status] if exc.__name__ in ("BufferFullError",) else []))

    def step(self, action):
        if self.done:
            raise EnvDone()
        if not self.action_space.contains(action):
            self._raise_for(abi.ST_ACTION_OUT_OF_SPACE if self._cand[3] > 0 else abi.ST_NO_POSSIBLE)
        self._batch.step(self._acts[int(action)])
        applied = (self._cand_transition(self._cand),) if int(action) == 1 else tuple()
        (reward, self._time, _, self._cand, self.terminated, self.truncated, status), obs = self._fetch()
        self._raise_for(status)
        self.done = self.terminated or self.truncated
        if status != abi.ST_STEP_FAILED:  # env.py:444-446: only successful results enter the history
            self._no_op_count += int(action == 0)
            self.history += (HistoryEntry(applied, True, "Success" if action == 1 else "NO OPERATION", self._time),)
        return obs, reward, self.terminated, self.truncated, self._get_info(action == 0)

    def op_log(self) -> np.ndarray:
        """Per-op (start, end) times so far -- what render() / the Gantt dashboard consume (8(f)1)."""
        return self._batch.get_op_log(0)

    def gantt_data(self) -> dict:
        """The data model of the reference's dashboard, DashboardDataMapper.map_states_to_gant_data(env.history)
        (rendering/gant_dashboard.py:414-429): {"schedules", "transports", "buffer"} -- rebuilt from the engine's
        state digest and transport-task log instead of a history of states (8(f)1).  Feed it to the reference's
        JobShopDashboard(data=..., ...) to get the Dash / Plotly view."""
        from .state_view import gantt_data
        return gantt_data(self.instance, self.state, self._batch.get_leg_log(0))

    def render(self, mode: str = "normal"):
        """Text Gantt of the same bars the reference's dashboard draws (machines, then transports); the Dash
        application itself is out of scope.  Returns the text."""
        data = self.gantt_data()
        rows = [(d["start"], d["end"], d["id"], d["job"], "") for d in data["schedules"]]
        rows += [(d["start"], d["end"] if d["end"] is not None else -1, d["id"], d["job"], d["meta_info"])
                 for d in data["transports"]]
        key = lambda r: (r[2][0] != "m", int(r[2][2:]), r[0])
        text = "\n".join(f"{a:>8} {b:>8}  {c:<6} {j:<6} {info}".rstrip() for a, b, c, j, info in sorted(rows, key=key))
        print(text)
        return text

    def close(self):
        if self._batch is not None:
            self._batch.close()
            self._batch = None
