"""Thin host binding of the CUDA engine (csrc/libjsl_b200.so) through its C ABI.

PyTorch is used for device memory and streams only: every array crossing the boundary is a raw
device pointer (``tensor.data_ptr()``).  There is NO CPU execution path -- if the shared library or a
CUDA device is missing, construction raises ``EngineUnavailable``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import abi
from .exceptions import EngineUnavailable, InvalidValue
from .lowered import LoweredInstance

import os

LIB_PATH = Path(os.environ.get("JSL_LIB", Path(__file__).resolve().parent / "csrc" / "libjsl_b200.so"))
LOG_LIB_PATH = Path(os.environ.get("JSL_LOG_LIB", Path(__file__).resolve().parent / "csrc" / "libjsl_b200_log.so"))
_libs: dict = {}

ABI_SYMBOLS = ("jsl_last_error", "jsl_abi_version", "jsl_instance_create", "jsl_instance_destroy", "jsl_batch_create",
               "jsl_batch_create_variants", "jsl_batch_load_variants", "jsl_batch_stage_variants", "jsl_batch_stage_variants_packed", "jsl_batch_commit_variants", "jsl_batch_check", "jsl_batch_randomize", "jsl_batch_get_variants", "jsl_batch_destroy", "jsl_batch_shape", "jsl_batch_state_layout", "jsl_batch_set_env_base", "jsl_batch_set_sample_tape", "jsl_batch_set_start_seeds",
               "jsl_reset", "jsl_step", "jsl_rollout", "jsl_get_state", "jsl_get_op_log", "jsl_get_leg_log", "jsl_launch_count")

RO_RESET_FIRST, RO_NO_WRITEBACK = 1, 2


def load_library(log: bool = False):
    """dlopen the engine; raises EngineUnavailable (never falls back to anything else).  `log`: the build that
    keeps the transport-task log (libjsl_b200_log.so, same ABI), used for batches created with record_ops."""
    if _libs.get(log) is not None:
        return _libs[log]
    path = LOG_LIB_PATH if log else LIB_PATH
    if not path.exists():
        raise EngineUnavailable(f"{path} is missing: build it with `python -m jobshoplab_b200.build` "
                                "(there is no CPU fallback)")
    L = C.CDLL(str(path))
    L.jsl_last_error.restype = C.c_char_p
    L.jsl_abi_version.restype = C.c_int
    L.jsl_instance_create.restype = C.c_void_p
    L.jsl_instance_create.argtypes = [C.POINTER(abi.InstanceDesc)]
    L.jsl_instance_destroy.argtypes = [C.c_void_p]
    L.jsl_batch_create.restype = C.c_void_p
    L.jsl_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                   C.POINTER(abi.EnvCfg)]
    L.jsl_batch_create_variants.restype = C.c_void_p
    L.jsl_batch_create_variants.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int, C.c_uint64, C.POINTER(abi.EnvCfg)]
    L.jsl_batch_load_variants.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.jsl_batch_stage_variants.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.jsl_batch_stage_variants_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.jsl_batch_commit_variants.argtypes = [C.c_void_p, C.c_void_p]
    L.jsl_batch_check.argtypes = [C.c_void_p]
    L.jsl_batch_randomize.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.jsl_batch_get_variants.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 5
    L.jsl_batch_destroy.argtypes = [C.c_void_p]
    L.jsl_batch_shape.argtypes = [C.c_void_p, C.POINTER(abi.Shape)]
    L.jsl_batch_state_layout.argtypes = [C.c_void_p, C.POINTER(abi.StateField), C.c_int]
    L.jsl_batch_set_env_base.argtypes = [C.c_void_p, C.c_int64]
    L.jsl_batch_set_sample_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.jsl_batch_set_start_seeds.argtypes = [C.c_void_p, C.c_void_p]
    L.jsl_reset.argtypes = [C.c_void_p, C.c_void_p, abi.StepOut, C.c_void_p]
    L.jsl_step.argtypes = [C.c_void_p, C.c_void_p, abi.StepOut, C.c_void_p]
    L.jsl_rollout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, abi.RolloutOut, C.c_void_p]
    L.jsl_get_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    L.jsl_get_op_log.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.jsl_get_leg_log.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    L.jsl_launch_count.restype = C.c_int64
    L.jsl_launch_count.argtypes = [C.c_void_p]
    if L.jsl_abi_version() != abi.ABI_VERSION:
        raise EngineUnavailable(f"{path.name} ABI version mismatch: rebuild it")
    _libs[log] = L
    return L


def _err(L) -> str:
    return (L.jsl_last_error() or b"").decode()


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise EngineUnavailable("no CUDA device: the jobshoplab_b200 engine runs on the GPU only")
    return torch


class Instance:
    """jsl_instance_t for one LoweredInstance."""

    def __init__(self, lowered: LoweredInstance, log: bool = False):
        self.L = load_library(log)
        self.lowered = lowered
        self._desc = abi.make_desc(lowered)
        self.h = self.L.jsl_instance_create(C.byref(self._desc))
        if not self.h:
            raise InvalidValue("instance", lowered.description, _err(self.L))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.jsl_instance_destroy(self.h)
            self.h = None


@dataclass
class StepResult:
    obs: "object"          # torch uint8 [B, obs_bytes] packed record (see Batch.unpack_obs)
    reward: "object"       # f64 [B]
    terminated: "object"   # u8 [B]
    truncated: "object"    # u8 [B]
    status: "object"       # u8 [B] JSL_ST_*
    time: "object"         # i32 [B]
    makespan: "object"     # i32 [B] (-1 unless terminated)
    cand: "object"         # i16 [B, 4]: kind, component, job, n_remaining of possible_transitions[0]
    ended: "object" = None  # u8 [B]: step(auto_reset=True): 0, or the JSL_ST_* the episode that was reset ended with


@dataclass
class RolloutResult:
    makespan: "object"
    n_steps: "object"
    ret: "object"
    status: "object"
    no_op_count: "object"


class Batch:
    """B envs on one CUDA device (jsl_batch_t).  Tensors are allocated once and reused."""

    def __init__(self, lowered, B: int, device: int = 0, seed: int = 0, variants=None, **cfg):
        torch = _torch()
        self.torch = torch
        self.L = load_library(log=bool(cfg.get("record_ops")))
        self.B = int(B)
        self.device = int(device)
        self.cfg = abi.make_cfg(**cfg)
        if isinstance(lowered, LoweredInstance):
            lowered = [lowered]
        self.lowered = list(lowered)
        self.instances = [Instance(li, log=bool(cfg.get("record_ops"))) for li in self.lowered]
        if variants is not None:
            om, od, lb, mat = (np.ascontiguousarray(a, dtype=np.int32) for a in variants)
            n_inst = om.shape[0]
            n_ops = self.lowered[0].n_ops
            if om.ndim != 2 or om.shape != (n_inst, n_ops) or od.shape != om.shape or lb.shape != (n_inst,) or mat.shape != (n_inst,):
                raise InvalidValue("variants", (om.shape, od.shape, lb.shape, mat.shape),
                                   f"expected op tables [n][{n_ops}] and bounds [n]")
            self.n_variants = int(n_inst)
            self.h = self.L.jsl_batch_create_variants(self.instances[0].h, n_inst, om.ctypes.data, od.ctypes.data,
                                                      lb.ctypes.data, mat.ctypes.data, self.B, self.device, seed,
                                                      C.byref(self.cfg))
        else:
            self.n_variants = len(self.instances)
            arr = (C.c_void_p * len(self.instances))(*[i.h for i in self.instances])
            self.h = self.L.jsl_batch_create(arr, len(self.instances), self.B, self.device, seed, C.byref(self.cfg))
        if not self.h:
            raise InvalidValue("batch", None, _err(self.L))
        self.shape = abi.Shape()
        self.L.jsl_batch_shape(self.h, C.byref(self.shape))
        dev = torch.device("cuda", self.device)
        ob = max(self.shape.obs_bytes, 16)
        # per-env scalars of a step arrive as one 32-byte jsl_step_record per env; the named fields are strided views
        # (records and observations share one allocation: a single-env batch is read back with one copy, fetch_host)
        RB = abi.STEP_RECORD_BYTES
        self._outbuf = torch.zeros(self.B * (RB + ob), dtype=torch.uint8, device=dev)
        self._host_out = None
        self.record = self._outbuf[: self.B * RB].view(self.B, RB)
        rec = self.record
        i32 = rec[:, 8:16].view(torch.int32)
        self.out = StepResult(
            obs=self._outbuf[self.B * RB:].view(self.B, ob),
            reward=rec[:, 0:8].view(torch.float64)[:, 0],
            terminated=rec[:, 24], truncated=rec[:, 25], status=rec[:, 26],
            time=i32[:, 0], makespan=i32[:, 1],
            cand=rec[:, 16:24].view(torch.int16), ended=rec[:, 27],
        )
        self.rout = RolloutResult(
            makespan=torch.zeros(self.B, dtype=torch.int32, device=dev),
            n_steps=torch.zeros(self.B, dtype=torch.int32, device=dev),
            ret=torch.zeros(self.B, dtype=torch.float64, device=dev),
            status=torch.zeros(self.B, dtype=torch.uint8, device=dev),
            no_op_count=torch.zeros(self.B, dtype=torch.int32, device=dev),
        )
        self._step_out = abi.StepOut(self.out.obs.data_ptr() if self.shape.obs_bytes else None, None, None, None, None,
                                     None, None, None, rec.data_ptr(), 0, 0)
        self._step_out_auto = abi.StepOut(self._step_out.obs, None, None, None, None, None, None, None, rec.data_ptr(),
                                          abi.STEP_AUTO_RESET, 0)
        r = self.rout
        self._roll_out = abi.RolloutOut(r.makespan.data_ptr(), r.n_steps.data_ptr(), r.ret.data_ptr(),
                                        r.status.data_ptr(), r.no_op_count.data_ptr())

    def close(self):
        if getattr(self, "h", None):
            self.L.jsl_batch_destroy(self.h)
            self.h = None

    __del__ = close

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise InvalidValue("engine call", rc, _err(self.L))

    def set_env_base(self, env_base: int):
        self._check(self.L.jsl_batch_set_env_base(self.h, int(env_base)))

    def _host_ptrs(self, arrays):
        ptrs = []
        n, N = int(self.n_variants), self.shape.n_ops
        for a, want in zip(arrays, (n * N, n * N, n, n)):
            if int(a.numel() if hasattr(a, "numel") else a.size) != want:
                raise InvalidValue("variant tables", tuple(a.shape), f"expected {want} int32 elements")
            if hasattr(a, "data_ptr"):
                assert not a.is_cuda and a.dtype == self.torch.int32 and a.is_contiguous()
                ptrs.append(a.data_ptr())
            else:
                assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
                ptrs.append(a.ctypes.data)
        return ptrs

    def load_variants(self, op_machine, op_duration, lower_bound, max_allowed_time):
        """Upload new operation tables for every variant from HOST int32 tensors/arrays (pinned torch
        tensors make the copies asynchronous on the current stream)."""
        ptrs = self._host_ptrs((op_machine, op_duration, lower_bound, max_allowed_time))
        self._check(self.L.jsl_batch_load_variants(self.h, *ptrs, self._stream()))

    def stage_variants(self, op_machine, op_duration, lower_bound, max_allowed_time):
        """First half of load_variants: host -> staging buffers on the current stream (e.g. a copy stream, while a
        rollout still runs on the tables in use)."""
        ptrs = self._host_ptrs((op_machine, op_duration, lower_bound, max_allowed_time))
        self._check(self.L.jsl_batch_stage_variants(self.h, *ptrs, self._stream()))

    def stage_variants_packed(self, op_machine_u8, op_duration_u16, lower_bound, max_allowed_time):
        """stage_variants with narrow HOST tables: uint8 machines, uint16 durations (3 instead of 8 bytes per operation
        over PCIe).  The caller guarantees the values fit (pack_variants checks)."""
        torch = self.torch
        n, N = int(self.n_variants), self.shape.n_ops
        for a, dt, want in ((op_machine_u8, torch.uint8, n * N), (op_duration_u16, torch.uint16, n * N),
                            (lower_bound, torch.int32, n), (max_allowed_time, torch.int32, n)):
            if a.is_cuda or a.dtype != dt or not a.is_contiguous() or a.numel() != want:
                raise InvalidValue("packed variant tables", (tuple(a.shape), a.dtype), f"expected {want} host elements of {dt}")
        self._check(self.L.jsl_batch_stage_variants_packed(self.h, op_machine_u8.data_ptr(), op_duration_u16.data_ptr(),
                                                           lower_bound.data_ptr(), max_allowed_time.data_ptr(), self._stream()))

    @staticmethod
    def pack_variants(op_machine, op_duration):
        """int32 tables -> (uint8 machines, uint16 durations) pinned host tensors for stage_variants_packed."""
        import torch
        om, od = torch.as_tensor(op_machine), torch.as_tensor(op_duration)
        if int(om.min()) < 0 or int(om.max()) > 255 or int(od.min()) < 0 or int(od.max()) > 65535:
            raise InvalidValue("variant tables", (int(om.max()), int(od.max())), "do not fit uint8 machines / uint16 durations")
        return om.to(torch.uint8).contiguous().pin_memory(), od.to(torch.uint16).contiguous().pin_memory()

    def check(self):
        """Raises if any uploaded table entry was out of range (validated on the device while packing; synchronises)."""
        self._check(self.L.jsl_batch_check(self.h))

    def commit_variants(self):
        """Second half: staged tables -> live tables on the current stream."""
        self._check(self.L.jsl_batch_commit_variants(self.h, self._stream()))

    def randomize(self, seed: int, min_duration: int | None = None, max_duration: int | None = None, mask=None):
        """InstanceRandomizer (compiler/manipulators.py:99-150) for the variants of this batch, on the device:
        per job a random permutation of its operations and durations uniform in [min_duration, max_duration)
        (defaults: min / max duration of the template instance, as the reference takes them from the instance
        it manipulates).  `mask`: uint8 CUDA tensor [n_variants] selecting the variants to redraw.  Applies
        from the next reset of the envs concerned."""
        d = np.asarray(self.lowered[0].op_duration)
        lo = int(d.min()) if min_duration is None else int(min_duration)
        hi = int(d.max()) if max_duration is None else int(max_duration)
        mp = None
        if mask is not None:
            assert mask.is_cuda and mask.dtype == self.torch.uint8 and mask.is_contiguous()
            mp = mask.data_ptr()
        rc = self.L.jsl_batch_randomize(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, lo, hi, mp, self._stream())
        if rc != 0:
            raise InvalidValue("randomize", (lo, hi), _err(self.L))

    def get_variants(self, first: int = 0, count: int | None = None):
        """(op_machine, op_tool, op_duration, lower_bound, max_allowed_time) of variants [first, first+count) as
        numpy int32 arrays (synchronises)."""
        n_var = int(self.n_variants)
        count = n_var - first if count is None else int(count)
        N = self.shape.n_ops
        om, ot, od = (np.zeros((count, N), np.int32) for _ in range(3))
        lb, mat = np.zeros(count, np.int32), np.zeros(count, np.int32)
        self._check(self.L.jsl_batch_get_variants(self.h, int(first), count, om.ctypes.data, ot.ctypes.data, od.ctypes.data,
                                                  lb.ctypes.data, mat.ctypes.data))
        return om, ot, od, lb, mat

    def set_sample_tape(self, tape=None):
        """Replay recorded StochasticTimeConfig.update() results (int32 CUDA tensor [B, L]) instead of drawing
        from the Philox streams; None switches back.  Applies from the next reset."""
        if tape is None:
            self._tape = None
            self._check(self.L.jsl_batch_set_sample_tape(self.h, None, 0))
            return
        assert tape.dtype == self.torch.int32 and tape.is_cuda and tape.dim() == 2 and tape.shape[0] == self.B
        self._tape = tape.contiguous()
        self._check(self.L.jsl_batch_set_sample_tape(self.h, self._tape.data_ptr(), self._tape.shape[1]))

    def set_start_seeds(self, seeds=None):
        """numpy start seeds of every time object (int32 CUDA tensor [B, n_time_objects], flat order of
        jobshoplab_b200.stochastic.flat_specs); None: drawn per env.  Applies from the next reset."""
        self._seeds = None if seeds is None else seeds.contiguous()
        if seeds is not None:
            assert seeds.dtype == self.torch.int32 and seeds.is_cuda and seeds.shape[0] == self.B
        self._check(self.L.jsl_batch_set_start_seeds(self.h, None if seeds is None else self._seeds.data_ptr()))

    def reset(self, mask=None) -> StepResult:
        ptr = None
        if mask is not None:
            assert mask.dtype == self.torch.uint8 and mask.is_cuda and mask.numel() == self.B
            ptr = mask.data_ptr()
        self._check(self.L.jsl_reset(self.h, ptr, self._step_out, self._stream()))
        return self.out

    def step(self, actions, auto_reset: bool = False) -> StepResult:
        """One env.step per env.  auto_reset (JSL_STEP_AUTO_RESET): envs whose episode ends in this step are reset in
        the same launch; `ended` tells which (0, or the status the episode ended with), reward / terminated /
        truncated / makespan are the finished step's, observation / time / status / cand the fresh episode's."""
        assert actions.dtype == self.torch.uint8 and actions.is_cuda and actions.numel() == self.B
        self._check(self.L.jsl_step(self.h, actions.data_ptr(), self._step_out_auto if auto_reset else self._step_out,
                                    self._stream()))
        return self.out

    def rollout(self, policy, tape=None, max_steps: int = 1 << 30, reset_first=False, writeback=True) -> RolloutResult:
        pol = abi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        ptr, stride = None, 0
        if tape is not None:
            assert tape.dtype == self.torch.uint8 and tape.is_cuda and tape.dim() == 2 and tape.shape[0] == self.B
            tape = tape.contiguous()
            ptr, stride = tape.data_ptr(), tape.shape[1]
        flags = (RO_RESET_FIRST if reset_first else 0) | (0 if writeback else RO_NO_WRITEBACK)
        self._check(self.L.jsl_rollout(self.h, pol, ptr, stride, int(min(max_steps, (1 << 31) - 1)), flags,
                                       self._roll_out, self._stream()))
        return self.rout

    def get_state(self, idx: int) -> np.ndarray:
        """Canonical int32 digest of env idx (synchronises)."""
        out = np.zeros(self.shape.digest_cap, dtype=np.int32)
        n = self.L.jsl_get_state(self.h, int(idx), out.ctypes.data, len(out))
        if n < 0:
            raise InvalidValue("jsl_get_state", n, _err(self.L))
        return out[:n].copy()

    def get_leg_log(self, idx: int) -> np.ndarray:
        """Transport tasks of env idx: int32 [n][5] = agv (< 0: invisible to env.history), job, start, end, route
        (jsl_get_leg_log; needs record_ops)."""
        cap = self.shape.n_ops + 2 * self.shape.n_jobs + 8
        out = np.zeros((cap, 5), dtype=np.int32)
        n = self.L.jsl_get_leg_log(self.h, int(idx), out.ctypes.data, cap)
        if n < 0:
            raise InvalidValue("leg log", n, _err(self.L))
        return out[:n]

    def get_op_log(self, idx: int) -> np.ndarray:
        out = np.zeros((self.shape.n_ops, 2), dtype=np.int32)
        self._check(self.L.jsl_get_op_log(self.h, int(idx), out.ctypes.data))
        return out

    @property
    def launches(self) -> int:
        return int(self.L.jsl_launch_count(self.h))

    def state_layout(self) -> dict:
        """{field: (offset, numpy dtype, count)} of the state image (= the OBS_STATE observation record)."""
        arr = (abi.StateField * 64)()
        n = self.L.jsl_batch_state_layout(self.h, arr, 64)
        if n < 0 or n > 64:
            raise InvalidValue("jsl_batch_state_layout", n, _err(self.L))
        out = {}
        for f in arr[:n]:
            dt = {(8, 1): np.float64, (4, 0): np.int32, (2, 0): np.uint16, (1, 0): np.uint8}[(f.elem_bytes, f.is_float)]
            out[f.name.decode()] = (int(f.offset), np.dtype(dt), int(f.count))
        return out

    def unpack_state(self, rec=None) -> dict:
        """Named zero-copy views into a [B, state_bytes] uint8 tensor of state images (the OBS_STATE observation):
        scalars as [B], arrays as [B, 32 * lane words] (entries past n_jobs / n_machines / n_transports are padding).
        j_* per job, m_* per machine, t_* per transport; meanings: csrc/jsl_engine.cuh EnvS."""
        torch = self.torch
        rec = self.out.obs if rec is None else rec
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.uint16): torch.int16,
               np.dtype(np.uint8): torch.uint8}
        views = {}
        for name, (off, dt, cnt) in self.state_layout().items():
            v = rec[:, off: off + dt.itemsize * cnt].view(tdt[dt])
            views[name] = v[:, 0] if cnt == 1 else v
        return views

    # ---- single-env read-back ---------------------------------------------------------------------------
    def fetch_host(self, idx: int = 0):
        """(jsl_step_record bytes, observation bytes) of env `idx` as numpy uint8 views of a pinned staging buffer:
        one device->host copy for a single-env batch, two otherwise, one stream synchronisation.  The views are
        overwritten by the next call."""
        torch, RB = self.torch, abi.STEP_RECORD_BYTES
        ob = self.out.obs.shape[1]
        if self._host_out is None:
            self._host_out = torch.empty(RB + ob, dtype=torch.uint8, pin_memory=True)
            self._host_np = self._host_out.numpy()
        if self.B == 1:
            self._host_out.copy_(self._outbuf, non_blocking=True)
        else:
            self._host_out[:RB].copy_(self.record[idx], non_blocking=True)
            self._host_out[RB:].copy_(self.out.obs[idx], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_np[:RB], self._host_np[RB:]

    def fetch_host_all(self):
        """(jsl_step_records [B, 32], observation records [B, obs_bytes]) of EVERY env as numpy uint8 views of a pinned
        staging buffer: one device->host copy and one stream synchronisation per call (what a host-side training loop
        -- vec_env.SB3VecEnv -- pays per step).  The views are overwritten by the next call."""
        torch, RB = self.torch, abi.STEP_RECORD_BYTES
        ob = self.out.obs.shape[1]
        if getattr(self, "_host_all", None) is None:
            self._host_all = torch.empty(self.B * (RB + ob), dtype=torch.uint8, pin_memory=True)
            self._host_all_np = self._host_all.numpy()
        self._host_all.copy_(self._outbuf, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_all_np[: self.B * RB].reshape(self.B, RB), self._host_all_np[self.B * RB:].reshape(self.B, ob)

    def unpack_obs_host(self, raw: np.ndarray) -> dict:
        """unpack_obs on the host: numpy views of one env's observation bytes `raw` [obs_bytes], or of all of them
        [B, obs_bytes] (then every entry has the leading batch axis of unpack_obs)."""
        J, M, N = self.shape.n_jobs, self.shape.n_machines, self.shape.n_ops
        kind = self.cfg.obs_kind
        lead = raw.shape[:-1]
        if kind == abi.OBS_STATE:
            t = self.unpack_state(self.torch.from_numpy(raw if lead else raw[None]))
            return {k: (v if lead else v[0]).numpy() for k, v in t.items()}
        if kind == abi.OBS_TASSEL:
            f = raw[..., : 28 * J].view(np.float32).reshape(lead + (7, J))
            keys = ("allocatable", "left_over_time", "percent_finished", "total_completion",
                    "time_until_next_machine_is_free", "idle_since_last_op", "cum_idle_time")
            return {k: f[..., i, :] for i, k in enumerate(keys)}
        if kind == abi.OBS_OPERATION_ARRAY:
            f = raw[..., : 16 + 4 * N + 4 * J].view(np.float32)
            return {"operation_state": f[..., None, 4: 4 + N], "job_locations": f[..., None, 4 + N: 4 + N + J],
                    "current_transition": f[..., 0:3]}
        o = 16
        f = raw[..., :16].view(np.float32)
        jprog = raw[..., o: o + 4 * J].view(np.int32); o += 4 * J
        mprog = raw[..., o: o + 4 * M].view(np.int32); o += 4 * M
        jrun = raw[..., o: o + J].view(np.int8); o += J
        avail = raw[..., o: o + J].view(np.int8); o += J
        mrun = raw[..., o: o + M].view(np.int8); o += M
        execd = raw[..., o: o + J * M].view(np.int8).reshape(lead + (J, M))
        return {"job_running": jrun, "job_executed_on_machine": execd, "job_progression": jprog,
                "machine_running": mrun, "machine_progression": mprog, "available_jobs": avail,
                "current_time": f[..., 0:1], "current_transition": f[..., 1:4]}

    # ---- observation record -> named views (zero copy) -------------------------------------------------
    def unpack_obs(self, obs=None) -> dict:
        """Views into the packed observation record, keyed like the reference's observation dict
        (jobshoplab/env/factories/observations.py:388-409, :507)."""
        torch = self.torch
        obs = self.out.obs if obs is None else obs
        J, M, N = self.shape.n_jobs, self.shape.n_machines, self.shape.n_ops
        if self.cfg.obs_kind == abi.OBS_STATE:
            return self.unpack_state(obs)
        if self.cfg.obs_kind == abi.OBS_TASSEL:
            f = obs[:, : 28 * J].view(torch.float32).view(-1, 7, J)
            keys = ("allocatable", "left_over_time", "percent_finished", "total_completion",
                    "time_until_next_machine_is_free", "idle_since_last_op", "cum_idle_time")
            return {k: f[:, i] for i, k in enumerate(keys)}
        if self.cfg.obs_kind == abi.OBS_OPERATION_ARRAY:
            f = obs[:, : 16 + 4 * N + 4 * J].view(torch.float32)
            return {"operation_state": f[:, 4: 4 + N].unsqueeze(1), "job_locations": f[:, 4 + N: 4 + N + J].unsqueeze(1),
                    "current_transition": f[:, 0:3]}
        o = 16
        f = obs[:, :16].view(torch.float32)
        jprog = obs[:, o: o + 4 * J].view(torch.int32); o += 4 * J
        mprog = obs[:, o: o + 4 * M].view(torch.int32); o += 4 * M
        jrun = obs[:, o: o + J].view(torch.int8); o += J
        avail = obs[:, o: o + J].view(torch.int8); o += J
        mrun = obs[:, o: o + M].view(torch.int8); o += M
        execd = obs[:, o: o + J * M].view(torch.int8).view(-1, J, M)
        return {"job_running": jrun, "job_executed_on_machine": execd, "job_progression": jprog,
                "machine_running": mrun, "machine_progression": mprog, "available_jobs": avail,
                "current_time": f[:, 0:1], "current_transition": f[:, 1:4]}
