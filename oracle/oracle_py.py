"""ctypes wrapper of the CPU oracle (oracle/jsl_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(_HERE.parent))
from jobshoplab_b200 import abi  # noqa: E402  (struct definitions only)

LIB_PATH = _HERE / "_build" / "libjsl_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "jsl_oracle.c"
    hdr = _HERE.parent / "include" / "jobshoplab_b200.h"
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= max(src.stat().st_mtime, hdr.stat().st_mtime):
        return LIB_PATH
    LIB_PATH.parent.mkdir(exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(LIB_PATH),
                           str(src), "-lpthread", "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.jslo_create.restype = C.c_void_p
        L.jslo_create.argtypes = [C.POINTER(abi.InstanceDesc), C.POINTER(abi.EnvCfg)]
        L.jslo_destroy.argtypes = [C.c_void_p]
        L.jslo_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for name in ("jslo_reset", "jslo_time", "jslo_terminated", "jslo_truncated", "jslo_status",
                     "jslo_n_steps", "jslo_noop_count", "jslo_n_possible", "jslo_obs_bytes",
                     "jslo_obs_oparray_bytes", "jslo_obs_tassel_bytes", "jslo_tape_pos"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.jslo_step.argtypes = [C.c_void_p, C.c_int]
        L.jslo_reward.argtypes = [C.c_void_p]
        L.jslo_reward.restype = C.c_double
        L.jslo_return.argtypes = [C.c_void_p]
        L.jslo_return.restype = C.c_double
        L.jslo_cand.argtypes = [C.c_void_p, C.c_void_p]
        L.jslo_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jslo_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jslo_obs_oparray.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jslo_obs_tassel.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jslo_random_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.jslo_policy_action.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
        L.jslo_rollout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        L.jslo_rollout_batch.restype = C.c_int64
        L.jslo_rollout_batch.argtypes = [C.POINTER(abi.InstanceDesc), C.c_int, C.POINTER(abi.EnvCfg),
                                         C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleEnv:
    """One env of the CPU oracle; mirrors JobShopLabEnv's step/reset surface."""

    def __init__(self, lowered, **cfg):
        self.L = lib()
        self.li = lowered
        self._desc = abi.make_desc(lowered)
        self._cfg = abi.make_cfg(**cfg)
        self.h = self.L.jslo_create(C.byref(self._desc), C.byref(self._cfg))
        self._tape = None
        self._dcap = 16 + 3 * lowered.n_jobs + 3 * lowered.n_ops + lowered.n_machines * (6 + 0) + \
            3 * lowered.n_jobs + 11 * lowered.n_transports + lowered.n_sbuf + lowered.n_jobs + \
            3 * (len(lowered.m_outage_freq) + lowered.n_transports * len(lowered.t_outage_freq)) + \
            1 + 4 * (lowered.n_jobs + lowered.n_transports * lowered.n_jobs) + 64
        self.reset()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.jslo_destroy(self.h)
            self.h = None

    def set_tape(self, samples):
        self._tape = np.ascontiguousarray(samples, dtype=np.int32)
        self.L.jslo_set_tape(self.h, self._tape.ctypes.data, len(self._tape))

    def reset(self):
        return self.L.jslo_reset(self.h)

    def step(self, a: int):
        return self.L.jslo_step(self.h, int(a))

    @property
    def time(self): return self.L.jslo_time(self.h)
    @property
    def reward(self): return self.L.jslo_reward(self.h)
    @property
    def ret(self): return self.L.jslo_return(self.h)
    @property
    def terminated(self): return bool(self.L.jslo_terminated(self.h))
    @property
    def truncated(self): return bool(self.L.jslo_truncated(self.h))
    @property
    def done(self): return self.terminated or self.truncated
    @property
    def status(self): return self.L.jslo_status(self.h)
    @property
    def n_steps(self): return self.L.jslo_n_steps(self.h)
    @property
    def noop_count(self): return self.L.jslo_noop_count(self.h)
    @property
    def n_possible(self): return self.L.jslo_n_possible(self.h)

    def cand(self):
        out = np.zeros(4, dtype=np.int32)
        self.L.jslo_cand(self.h, out.ctypes.data)
        return out

    def digest(self) -> np.ndarray:
        out = np.zeros(self._dcap, dtype=np.int32)
        n = self.L.jslo_digest(self.h, out.ctypes.data, len(out))
        if n > len(out):
            out = np.zeros(n, dtype=np.int32)
            n = self.L.jslo_digest(self.h, out.ctypes.data, len(out))
        return out[:n].copy()

    def obs(self) -> bytes:
        n = self.L.jslo_obs_bytes(self.h)
        out = np.zeros(n, dtype=np.uint8)
        r = self.L.jslo_obs(self.h, out.ctypes.data, n)
        if r < 0:
            raise RuntimeError(f"oracle obs raised status {-r}")
        return out.tobytes()

    def obs_oparray(self) -> bytes:
        n = self.L.jslo_obs_oparray_bytes(self.h)
        out = np.zeros(n, dtype=np.uint8)
        r = self.L.jslo_obs_oparray(self.h, out.ctypes.data, n)
        if r < 0:
            raise RuntimeError(f"oracle obs raised status {-r}")
        return out.tobytes()

    def obs_tassel(self) -> bytes:
        n = self.L.jslo_obs_tassel_bytes(self.h)
        out = np.zeros(n, dtype=np.uint8)
        self.L.jslo_obs_tassel(self.h, out.ctypes.data, n)
        return out.tobytes()

    def policy_action(self, policy: int, seed: int = 0, env_idx: int = 0) -> int:
        return self.L.jslo_policy_action(self.h, policy, seed, env_idx)

    def rollout(self, policy: int, tape=None, max_steps=1 << 30, seed=0, env_idx=0) -> int:
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint8)
            return self.L.jslo_rollout(self.h, policy, tape.ctypes.data, len(tape), max_steps, seed, env_idx)
        return self.L.jslo_rollout(self.h, policy, None, 0, max_steps, seed, env_idx)


def rollout_batch(lowered_list, B, policy, seed=0, max_steps=1 << 30, n_threads=None, env_base=0, **cfg):
    """B independent episodes on host threads; returns dict of arrays + total steps."""
    L = lib()
    descs = (abi.InstanceDesc * len(lowered_list))(*[abi.make_desc(li) for li in lowered_list])
    c = abi.make_cfg(**cfg)
    mk = np.zeros(B, dtype=np.int32); ns = np.zeros(B, dtype=np.int32)
    ret = np.zeros(B, dtype=np.float64); st = np.zeros(B, dtype=np.uint8)
    total = L.jslo_rollout_batch(descs, len(lowered_list), C.byref(c), B, policy, seed, max_steps,
                                 n_threads or os.cpu_count() or 1, env_base,
                                 mk.ctypes.data, ns.ctypes.data, ret.ctypes.data, st.ctypes.data)
    return {"makespan": mk, "n_steps": ns, "ret": ret, "status": st, "total_steps": int(total)}
