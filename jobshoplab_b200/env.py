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

from . import abi, compiler as _compiler
from .config import Config
from .engine import Batch
from .exceptions import (STATUS_TO_EXCEPTION, ConfigurationError, EnvDone, InvalidValue)
from .lowered import LoweredInstance
from .state_view import parse_digest

_OBS_KINDS = {"BinaryActionObservationFactory": abi.OBS_BINARY_ACTION,
              "BinaryOperationArrayObservation": abi.OBS_OPERATION_ARRAY,
              "TasselJsspObservation": abi.OBS_TASSEL}


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


def _lower(config: Config, compiler) -> LoweredInstance:
    if isinstance(compiler, LoweredInstance):
        return compiler
    if compiler is None:
        compiler = _compiler.Compiler(config)
    out = compiler.compile()
    if not isinstance(out, LoweredInstance):
        raise InvalidValue("compiler", type(out), "compile() must return a LoweredInstance "
                           "(use jobshoplab_b200.compiler.Compiler)")
    return out


class BatchedJobShopLabEnv:
    """B envs of one instance (or of same-shaped instances) on one GPU."""

    def __init__(self, config: Config | None = None, num_envs: int = 1, compiler=None, device: int = 0, seed: int = 0,
                 instances=None, variants=None, record_ops: bool = False, auto_reset: bool = False):
        self.config = config or Config()
        lowered = instances if instances is not None else _lower(self.config, compiler)
        # compiler.manipulators = [InstanceRandomizer]: the reference recompiles -- and so re-randomises -- the
        # instance on every env.reset (env.py:491).  Batched: every env owns a variant of the operation
        # tables, redrawn on the device whenever that env is reset (engine.Batch.randomize).
        self.randomize = "InstanceRandomizer" in [m for m in self.config.compiler.manipulators if m]
        if self.randomize and variants is None and instances is None:
            li = lowered
            tile = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.int32), (num_envs, len(a))))
            variants = (tile(li.op_machine), tile(li.op_duration), np.full(num_envs, li.lower_bound, np.int32),
                        np.full(num_envs, li.max_allowed_time, np.int32))
        self.batch = Batch(lowered, num_envs, device=device, seed=seed, variants=variants, record_ops=record_ops,
                           **_cfg_kwargs(self.config))
        self.instance = self.batch.lowered[0]
        self.num_envs = num_envs
        self.auto_reset = auto_reset
        self.action_space = Discrete(2)
        self._seed, self._epoch = int(seed), 0
        d = np.asarray(self.instance.op_duration)
        self._dur_range = (int(d.min()), int(d.max()))  # of the instance handed to the manipulator

    def _redraw(self, mask=None):
        self._epoch += 1
        self.batch.randomize(seed=(self._seed * 0x9E3779B97F4A7C15 + self._epoch) & 0xFFFFFFFFFFFFFFFF,
                             min_duration=self._dur_range[0], max_duration=self._dur_range[1], mask=mask)

    def reset(self, mask=None):
        if self.randomize:
            self._redraw(mask)
        out = self.batch.reset(mask)
        return self.batch.unpack_obs(out.obs), out

    def step(self, actions):
        """actions: uint8 CUDA tensor [B].  Returns (obs dict of views, reward, terminated, truncated, out)."""
        out = self.batch.step(actions)
        if self.auto_reset:
            done = (out.terminated | out.truncated)
            if bool(done.any()):
                final = {k: v.clone() for k, v in (("reward", out.reward), ("terminated", out.terminated),
                                                   ("truncated", out.truncated), ("makespan", out.makespan))}
                if self.randomize:
                    self._redraw(done)
                self.batch.reset(done)
                out.reward.copy_(final["reward"]); out.terminated.copy_(final["terminated"])
                out.truncated.copy_(final["truncated"]); out.makespan.copy_(final["makespan"])
        return self.batch.unpack_obs(out.obs), out.reward, out.terminated, out.truncated, out

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


class JobShopLabEnv:
    """Single-env drop-in (env.py:293-618).  `compiler` may be a jobshoplab_b200.compiler.Compiler, a
    LoweredInstance, or None (built from config.compiler like the reference)."""

    metadata = {"render_modes": ["normal", "dashboard", "debug"]}

    def __init__(self, config: Config | None = None, seed=None, compiler=None, observation_factory=None,
                 reward_factory=None, middleware=None, action_factory=None, render_backend=None, loglevel=None,
                 device: int = 0):
        for name, arg in (("observation_factory", observation_factory), ("reward_factory", reward_factory),
                          ("middleware", middleware), ("action_factory", action_factory)):
            if arg is not None and not isinstance(arg, str):
                raise ConfigurationError(name, arg, "the CUDA engine implements the reference's stock classes; "
                                         "select them by name through the config")
        self.config = config or Config()
        if isinstance(observation_factory, str):
            self.config.env.observation_factory = observation_factory
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

    def _obs_host(self):
        views = self._batch.unpack_obs(self._batch.out.obs)
        host = {k: v[0].cpu().numpy() for k, v in views.items()}
        if self._batch.cfg.obs_kind == abi.OBS_BINARY_ACTION:
            for k in ("job_running", "machine_running", "available_jobs"):
                host[k] = tuple(bool(x) for x in host[k])
            host["job_executed_on_machine"] = tuple(tuple(bool(x) for x in row) for row in host["job_executed_on_machine"])
            host["job_progression"] = tuple(int(x) for x in host["job_progression"])
            host["machine_progression"] = tuple(int(x) for x in host["machine_progression"])
        return host

    def _get_info(self, no_op: bool) -> dict:
        return {"no_op": no_op, "terminated": self.terminated, "truncated": self.truncated,
                "makespan": int(self._batch.out.time[0]) if self.terminated else None,
                "no_op_count": self._no_op_count}

    def reset(self, seed=None):
        self.episode_counter += 1
        if seed is None:
            seed = self.seed
        self.seed = seed + self.episode_counter if seed is not None else None
        self.instance = _lower(self.config, self._compiler)  # the reference recompiles on every reset (env.py:491)
        if self._batch is not None:
            self._batch.close()
        self._batch = Batch(self.instance, 1, device=self._device, seed=self.seed or 0, record_ops=True,
                            **_cfg_kwargs(self.config))
        self.lower_bound = self.instance.lower_bound
        self.max_allowed_time = self.instance.max_allowed_time
        self.action_space = Discrete(2)
        self.observation_space = None
        out = self._batch.reset()
        self._raise_for(int(out.status[0]))
        self.terminated = self.truncated = self.done = False
        self._no_op_count = 0
        self._acts = self._batch.torch.zeros(1, dtype=self._batch.torch.uint8, device=f"cuda:{self._device}")
        return self._obs_host(), {"env_seed": self.seed}

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
            status = abi.ST_ACTION_OUT_OF_SPACE if self._batch.out.cand[0, 3] > 0 else abi.ST_NO_POSSIBLE
            self._raise_for(status)
        self._acts.fill_(int(action))
        out = self._batch.step(self._acts)
        status = int(out.status[0])
        self._raise_for(status)
        self.terminated = bool(out.terminated[0])
        self.truncated = bool(out.truncated[0])
        self.done = self.terminated or self.truncated
        if status != abi.ST_STEP_FAILED:
            self._no_op_count += int(action == 0)
        return self._obs_host(), float(out.reward[0]), self.terminated, self.truncated, self._get_info(action == 0)

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
