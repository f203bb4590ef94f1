"""ctypes wrapper of tests/hostsim/jsl_hostsim.cpp -- the CUDA engine source compiled for the CPU with
loop primitives.  DEBUG / TEST TOOLING ONLY (see the .cpp header); never imported by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from jobshoplab_b200 import abi

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libjsl_hostsim.so"
_lib = None


def build(force=False):
    srcs = [HERE / "jsl_hostsim.cpp", ROOT / "jobshoplab_b200" / "csrc" / "jsl_engine.cuh",
            ROOT / "jobshoplab_b200" / "csrc" / "jsl_tobj.hpp", ROOT / "jobshoplab_b200" / "csrc" / "jsl_randomize.cuh",
            ROOT / "include" / "jobshoplab_b200.h"]
    if not force and LIB.exists() and LIB.stat().st_mtime >= max(s.stat().st_mtime for s in srcs):
        return LIB
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           "-Wno-unknown-pragmas", "-DJSL_LEG_LOG", "-o", str(LIB), str(srcs[0])])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.jsh_create.restype = C.c_void_p
        L.jsh_create.argtypes = [C.POINTER(abi.InstanceDesc), C.POINTER(abi.EnvCfg), C.c_uint64]
        L.jsh_destroy.argtypes = [C.c_void_p]
        L.jsh_reset.argtypes = [C.c_void_p]
        L.jsh_step.argtypes = [C.c_void_p, C.c_int]
        L.jsh_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jsh_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.jsh_rollout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64]
        L.jsh_scalars.argtypes = [C.c_void_p, C.c_void_p]
        L.jsh_reward.argtypes = [C.c_void_p]
        L.jsh_reward.restype = C.c_double
        L.jsh_return.argtypes = [C.c_void_p]
        L.jsh_return.restype = C.c_double
        L.jsh_state_bytes.argtypes = [C.c_void_p]
        L.jsh_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jsh_set_env_id.argtypes = [C.c_void_p, C.c_uint64]
        L.jsh_set_start_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jsh_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_float]
        _lib = L
    return _lib


def invariants():
    """(checks, violations) of the event-loop shortcut 'an occupied_till-driven transition is due at now =>
    the time machine would not move' (jsl_engine.cuh sm_step), accumulated by the CPU debug build."""
    o = (C.c_long * 2)()
    lib().jsh_invariants(o)
    return int(o[0]), int(o[1])


def randomize_variant(seed, variant, job_op_start, op_machine, op_tool, min_d, max_d):
    """jsl_randomize.cuh randomize_variant on the CPU: returns (op_machine, op_tool, op_duration, job_work, lb, mat)."""
    jos = np.ascontiguousarray(job_op_start, np.int32)
    mt = np.ascontiguousarray(np.asarray(op_machine, np.int32) | (np.asarray(op_tool, np.int32) << 16))
    N, J = len(mt), len(jos) - 1
    dur, work, lm = np.zeros(N, np.int32), np.zeros(J, np.int32), np.zeros(2, np.int32)
    L = lib()
    L.jsh_randomize_variant.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    L.jsh_randomize_variant(int(seed), int(variant), J, jos.ctypes.data, int(min_d), int(max_d), mt.ctypes.data,
                            dur.ctypes.data, work.ctypes.data, lm.ctypes.data)
    return mt & 0xffff, mt >> 16, dur, work, int(lm[0]), int(lm[1])


def fused():
    """How often each chain fusion of the general event loop fired so far (jsl_engine.cuh apply_machine /
    apply_transport): [IDLE->SETUP->WORKING, (unused), TRANSIT->OUTAGE->IDLE, IDLE->PICKUP->WAITING,
    PICKUP->WAITING->TRANSIT]."""
    o = (C.c_long * 5)()
    lib().jsh_fused(o)
    return list(o)


def set_fuse(mask: int):
    """Debug switch of the CPU build: bit i enables fusion i (-1: all, the kernels' behaviour; 0: one reference batch per
    loop iteration)."""
    lib().jsh_set_fuse(int(mask))


def fast_paths():
    """(classic_start hits, classic_complete hits, batches of classic instances that took the general loop)."""
    o = (C.c_long * 3)()
    lib().jsh_fast_paths(o)
    return int(o[0]), int(o[1]), int(o[2])


class HostSimEnv:
    def __init__(self, lowered, seed=0, **cfg):
        self.L = lib()
        self.li = lowered
        self._desc = abi.make_desc(lowered)
        cfg.setdefault("record_ops", True)
        self._cfg = abi.make_cfg(**cfg)
        self.h = self.L.jsh_create(C.byref(self._desc), C.byref(self._cfg), seed)
        if not self.h:
            raise RuntimeError("hostsim: unsupported shape")
        self._obs_n = 16 + 16 * (lowered.n_jobs * lowered.n_machines + 8 * (lowered.n_jobs + lowered.n_machines) + 4 * lowered.n_ops)
        self.reset()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.jsh_destroy(self.h)
            self.h = None

    def set_tape(self, samples):
        t = np.ascontiguousarray(samples, dtype=np.int32)
        self.L.jsh_set_tape(self.h, t.ctypes.data, len(t))

    def set_start_seeds(self, seeds):
        t = np.ascontiguousarray(seeds, dtype=np.int32)
        self.L.jsh_set_start_seeds(self.h, t.ctypes.data, len(t))

    def set_env_id(self, i): self.L.jsh_set_env_id(self.h, int(i))

    def reset(self): return self.L.jsh_reset(self.h)
    def step(self, a): return self.L.jsh_step(self.h, int(a))

    def _sc(self):
        o = np.zeros(8, dtype=np.int32)
        self.L.jsh_scalars(self.h, o.ctypes.data)
        return o

    time = property(lambda s: int(s._sc()[0]))
    terminated = property(lambda s: bool(s._sc()[1]))
    truncated = property(lambda s: bool(s._sc()[2]))
    status = property(lambda s: int(s._sc()[3]))
    n_steps = property(lambda s: int(s._sc()[4]))
    noop_count = property(lambda s: int(s._sc()[5]))
    n_possible = property(lambda s: int(s._sc()[6]))
    reward = property(lambda s: s.L.jsh_reward(s.h))
    ret = property(lambda s: s.L.jsh_return(s.h))
    done = property(lambda s: s.terminated or s.truncated)

    def digest(self):
        out = np.zeros(200000, dtype=np.int32)
        n = self.L.jsh_digest(self.h, out.ctypes.data, len(out))
        return out[:n].copy()

    def obs(self) -> bytes:
        out = np.zeros(self._obs_n, dtype=np.uint8)
        n = self.L.jsh_obs(self.h, out.ctypes.data)
        if n < 0:
            raise RuntimeError(f"obs raised status {-n}")
        return out[:n].tobytes()

    def leg_log(self):
        cap = self.li.n_ops + 2 * self.li.n_jobs + 8
        out = np.zeros((cap, 5), dtype=np.int32)
        self.L.jsh_leg_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        n = self.L.jsh_leg_log(self.h, out.ctypes.data, cap)
        assert n >= 0, n
        return out[:n]

    def rollout(self, policy, tape=None, max_steps=1 << 30, env_idx=0):
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint8)
            return self.L.jsh_rollout(self.h, policy, tape.ctypes.data, len(tape), max_steps, env_idx)
        return self.L.jsh_rollout(self.h, policy, None, 0, max_steps, env_idx)
