"""ctypes mirror of include/jobshoplab_b200.h (structs, constants) -- no compute here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lowered import LoweredInstance, TimeSpec

ABI_VERSION = 2

(ST_OK, ST_DONE, ST_TRUNCATED, ST_STEP_FAILED, ST_BUFFER_FULL, ST_UNSUCCESSFUL, ST_NO_POSSIBLE,
 ST_ENV_DONE, ST_ACTION_OUT_OF_SPACE, ST_OTHER_INVALID) = range(10)
STATUS_NAMES = ("OK", "DONE", "TRUNCATED", "STEP_FAILED", "BUFFER_FULL", "UNSUCCESSFUL",
                "NO_POSSIBLE", "ENV_DONE", "ACTION_OUT_OF_SPACE", "OTHER_INVALID")
POLICY_FIFO, POLICY_SPT, POLICY_MWKR, POLICY_RANDOM, POLICY_TAPE = range(5)
POLICIES = {"fifo": POLICY_FIFO, "spt": POLICY_SPT, "mwkr": POLICY_MWKR, "random": POLICY_RANDOM,
            "tape": POLICY_TAPE}
M_IDLE, M_SETUP, M_WORKING, M_OUTAGE = range(4)                                  # JSL_M_*
T_IDLE, T_WORKING, T_PICKUP, T_TRANSIT, T_OUTAGE, T_WAITINGPICKUP = range(6)        # JSL_T_*
OBS_NONE, OBS_BINARY_ACTION, OBS_OPERATION_ARRAY, OBS_TASSEL, OBS_STATE = 0, 1, 2, 3, 4

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


class TimeSpecC(C.Structure):
    _fields_ = [("kind", _i32p), ("base", _i32p), ("param", _f32p)]


class InstanceDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_jobs", C.c_int32), ("n_machines", C.c_int32), ("n_transports", C.c_int32),
        ("n_sbuf", C.c_int32), ("n_tools", C.c_int32), ("n_ops", C.c_int32),
        ("job_op_start", _i32p), ("op_machine", _i32p), ("op_tool", _i32p), ("op_duration", _i32p),
        ("pre_type", _i32p), ("pre_cap", _i32p), ("post_type", _i32p), ("post_cap", _i32p),
        ("setup_time", _i32p), ("default_tool", C.c_int32),
        ("sbuf_type", _i32p), ("sbuf_cap", _i32p), ("sbuf_role", _i32p), ("buf_ref_id", _i32p),
        ("travel_time", _i32p), ("agv_init_place", _i32p), ("job_init_buf", _i32p),
        ("start_time", C.c_int32),
        ("m_outage_start", _i32p), ("m_outage_freq", _i32p), ("m_outage_dur", _i32p),
        ("n_t_outages", C.c_int32), ("t_outage_freq", _i32p), ("t_outage_dur", _i32p),
        ("op_spec", TimeSpecC), ("setup_spec", TimeSpecC), ("travel_spec", TimeSpecC),
        ("m_ofreq_spec", TimeSpecC), ("m_odur_spec", TimeSpecC), ("t_ofreq_spec", TimeSpecC),
        ("t_odur_spec", TimeSpecC),
        ("sample_table", _i32p), ("sample_row", _i32p), ("n_sample_rows", C.c_int32), ("sample_depth", C.c_int32),
        ("lower_bound", C.c_int32), ("max_allowed_time", C.c_int32),
    ]


class EnvCfg(C.Structure):
    _fields_ = [
        ("allow_early_transport", C.c_int32), ("truncation_joker", C.c_int32),
        ("truncation_active", C.c_int32), ("obs_kind", C.c_int32),
        ("sparse_bias", C.c_double), ("dense_bias", C.c_double), ("truncation_bias", C.c_double),
        ("record_ops", C.c_int32), ("reserved", C.c_int32),
    ]


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_jobs", "n_machines", "n_transports", "n_sbuf", "n_ops",
                                         "n_tools", "obs_bytes", "state_bytes", "digest_cap")]


class StepOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("status", C.c_void_p), ("time", C.c_void_p),
                ("makespan", C.c_void_p), ("cand", C.c_void_p), ("record", C.c_void_p), ("flags", C.c_uint32),
                ("reserved", C.c_uint32)]


STEP_RECORD_BYTES = 32  # jsl_step_record: f64 reward | i32 time | i32 makespan | i16 cand[4] | u8 terminated, truncated, status, ended
STEP_AUTO_RESET = 1     # jsl_step_out.flags: JSL_STEP_AUTO_RESET


class StateField(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("offset", C.c_int32), ("elem_bytes", C.c_int32), ("count", C.c_int32),
                ("is_float", C.c_int32)]


class RolloutOut(C.Structure):
    _fields_ = [("makespan", C.c_void_p), ("n_steps", C.c_void_p), ("ret", C.c_void_p),
                ("status", C.c_void_p), ("no_op_count", C.c_void_p)]


def _p(a: np.ndarray | None, ty=_i32p):
    if a is None or a.size == 0:
        return ty()
    return a.ctypes.data_as(ty)


def _spec(sp: TimeSpec) -> TimeSpecC:
    if sp.kind is None:
        return TimeSpecC()
    return TimeSpecC(_p(sp.kind), _p(sp.base), _p(sp.param, _f32p))


def make_desc(li: LoweredInstance) -> InstanceDesc:
    """Borrowing view: `li` must outlive the returned struct."""
    d = InstanceDesc()
    d.abi_version = ABI_VERSION
    d.n_jobs, d.n_machines, d.n_transports = li.n_jobs, li.n_machines, li.n_transports
    d.n_sbuf, d.n_tools, d.n_ops = li.n_sbuf, li.n_tools, li.n_ops
    for name in ("job_op_start", "op_machine", "op_tool", "op_duration", "pre_type", "pre_cap",
                 "post_type", "post_cap", "setup_time", "sbuf_type", "sbuf_cap", "sbuf_role",
                 "buf_ref_id", "travel_time", "agv_init_place", "job_init_buf", "m_outage_start",
                 "m_outage_freq", "m_outage_dur", "t_outage_freq", "t_outage_dur"):
        setattr(d, name, _p(getattr(li, name)))
    d.default_tool = li.default_tool
    d.start_time = li.start_time
    d.n_t_outages = len(li.t_outage_freq)
    for name in ("op_spec", "setup_spec", "travel_spec", "m_ofreq_spec", "m_odur_spec",
                 "t_ofreq_spec", "t_odur_spec"):
        setattr(d, name, _spec(getattr(li, name)))
    if not li.is_deterministic:
        if getattr(li, "_sample_tables", None) is None:
            from .stochastic import sample_tables
            li._sample_tables = sample_tables(li)
        table, row = li._sample_tables
        d.sample_table, d.sample_row = _p(table), _p(row)
        d.n_sample_rows, d.sample_depth = table.shape
    d.lower_bound, d.max_allowed_time = li.lower_bound, li.max_allowed_time
    return d


def make_cfg(allow_early_transport=True, truncation_joker=5, truncation_active=False,
             obs_kind=OBS_BINARY_ACTION, sparse_bias=1.0, dense_bias=0.001, truncation_bias=-1.0,
             record_ops=False) -> EnvCfg:
    return EnvCfg(int(allow_early_transport), int(truncation_joker), int(truncation_active),
                  int(obs_kind), float(sparse_bias), float(dense_bias), float(truncation_bias),
                  int(record_ops), 0)
