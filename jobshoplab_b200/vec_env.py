"""Vector-env adapters over BatchedJobShopLabEnv.

The reference reserved jobshoplab/env/env_wrapper.py and tests/end_to_end_tests/test_vec_env_end_to_end.py for a
vectorised env and shipped neither; its notebooks train PPO (stable-baselines3) on the single env.  `SB3VecEnv`
speaks the stable_baselines3 ``VecEnv`` protocol -- ``reset() -> obs``, ``step_async(actions)``,
``step_wait() -> (obs, rewards, dones, infos)`` with SB3's auto-reset convention (a finished env is reset at once,
its last observation goes to ``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]`` tells a
truncation from a termination), ``get_attr / set_attr / env_method / env_is_wrapped / seed / close`` -- with numpy
observations, so PPO consumes B GPU envs like any other VecEnv.  If stable_baselines3 is importable the class derives
from its ``VecEnv``; it does not need it (neither SB3 nor gymnasium exist in the build container: spaces fall back to
the small stand-ins below, which carry the reference's shapes / dtypes / bounds, observations.py:388-409, :507).

Everything on the device stays in BatchedJobShopLabEnv; this module only moves the observation record and the
per-step scalars to the host once per step.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import abi
from .env import BatchedJobShopLabEnv

try:  # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:  # noqa: BLE001
    _VecEnvBase = object


from .spaces import Box, DictSpace, DiscreteSpace, _spaces, observation_space  # noqa: E402,F401


class SB3VecEnv(_VecEnvBase):
    """stable_baselines3-style VecEnv over B GPU envs (see the module docstring)."""

    def __init__(self, config=None, num_envs: int = 1, **kw):
        self.venv = kw.pop("venv", None) or BatchedJobShopLabEnv(config, num_envs=num_envs, auto_reset=False, **kw)
        b = self.venv.batch
        self.num_envs = self.venv.num_envs
        self.observation_space = observation_space(b.cfg.obs_kind, b.shape.n_jobs, b.shape.n_machines, b.shape.n_ops)
        self.action_space = _spaces.Discrete(2) if _spaces else DiscreteSpace(2)
        if _VecEnvBase is not object:  # stable_baselines3 installed: its own bookkeeping (reset_infos, _seeds, _options)
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)
        self.render_mode = None
        self.metadata = {"render_modes": []}
        self._torch = b.torch
        self._actions = self._torch.zeros(self.num_envs, dtype=self._torch.uint8, device=f"cuda:{b.device}")
        self._episode_return = np.zeros(self.num_envs, np.float64)
        self._episode_length = np.zeros(self.num_envs, np.int64)
        self._pending = False

    # ------------------------------------------------------------------ helpers
    def _host_obs(self, views) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((k, views[k].cpu().numpy().copy()) for k in self.observation_space.keys())

    def _fetch(self):
        """Step records and observations of all envs with ONE device->host copy (Batch.fetch_host_all):
        (records [B, 32] uint8 view, observation dict of fresh numpy arrays)."""
        rec, raw = self.venv.batch.fetch_host_all()
        views = self.venv.batch.unpack_obs_host(raw)
        return rec, OrderedDict((k, np.array(views[k])) for k in self.observation_space.keys())

    # ------------------------------------------------------------------ VecEnv protocol
    def reset(self):
        views, _ = self.venv.reset()
        self._episode_return[:] = 0
        self._episode_length[:] = 0
        return self._host_obs(views) if self.venv.observation_fn else self._fetch()[1]

    def step_async(self, actions):
        a = np.asarray(actions).astype(np.uint8).reshape(self.num_envs)
        self._actions.copy_(self._torch.from_numpy(a), non_blocking=True)
        self._out = self.venv.batch.step(self._actions)
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_wait() without step_async()"
        self._pending = False
        rec, obs = self._fetch()  # the only device->host traffic of the step (plus one more for the envs reset below)
        rewards = rec[:, 0:8].view(np.float64)[:, 0].astype(np.float32)
        terminated, truncated = rec[:, 24].astype(bool), rec[:, 25].astype(bool)
        makespan = rec[:, 12:16].view(np.int32)[:, 0].copy()
        status = rec[:, 26].copy()
        # an env that died with an escaped exception (status >= JSL_ST_STEP_FAILED: BufferFullError, ...) is over as
        # well: it is reported (infos[i]["status"], "TimeLimit.truncated" False) and reset like any finished env
        failed = status >= abi.ST_STEP_FAILED
        dones = terminated | truncated | failed
        self._episode_return += rewards
        self._episode_length += 1
        infos = [{} for _ in range(self.num_envs)]
        if dones.any():
            idx = np.nonzero(dones)[0]
            for i in idx:
                infos[i] = {"terminal_observation": OrderedDict((k, v[i].copy()) for k, v in obs.items()),
                            "TimeLimit.truncated": bool(truncated[i] and not terminated[i]),
                            "makespan": int(makespan[i]) if terminated[i] else None, "status": int(status[i]),
                            "episode": {"r": float(self._episode_return[i]), "l": int(self._episode_length[i])}}
            mask = self._torch.from_numpy(dones.astype(np.uint8)).to(self._actions.device)
            self.venv.reset(mask)  # (redraws the instances of those envs under InstanceRandomizer)
            fresh = self._fetch()[1]
            for k in obs:
                obs[k][idx] = fresh[k][idx]
            self._episode_return[idx] = 0
            self._episode_length[idx] = 0
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.venv.close()

    def seed(self, seed=None):
        return [None] * self.num_envs  # the engine is deterministic given the batch seed

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else indices

    def get_attr(self, attr_name, indices=None):
        src = self.venv if hasattr(self.venv, attr_name) else self.venv.batch
        return [getattr(src, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.venv, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(self.venv, method_name)(*args, **kwargs) for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None
