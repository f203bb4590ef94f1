"""Loader for tests/golden/*.npz (made by oracle/gen_golden.py from the unmodified reference)."""
from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from jobshoplab_b200 import lowered as L

GOLDEN = Path(__file__).resolve().parent / "golden"
GROUPS = ("classic", "solutions", "transport", "buffers", "features", "truncation", "oparray", "tassel", "big")

_P = np.uint64(1099511628211)
_H0 = np.uint64(1469598103934665603)
_POW = None


def _powers(n):
    global _POW
    if _POW is None or len(_POW) < n + 1:
        m = max(n + 1, 4096)
        p = np.empty(m, dtype=np.uint64)
        p[0] = 1
        with np.errstate(over="ignore"):
            np.multiply.accumulate(np.full(m - 1, _P, dtype=np.uint64), out=p[1:])
        _POW = p
    return _POW


def hash_i32(a) -> int:
    """h = H0; for x: h = h*P + (u32)x + 1 (mod 2^64) -- vectorised."""
    x = np.ascontiguousarray(a, dtype=np.int32).view(np.uint32).astype(np.uint64) + np.uint64(1)
    n = len(x)
    pw = _powers(n)
    with np.errstate(over="ignore"):
        return int(_H0 * pw[n] + np.sum(x * pw[n - 1::-1][:n] if n else x, dtype=np.uint64))


def hash_bytes(b: bytes) -> int:
    pad = b + b"\0" * (-len(b) % 4)
    return hash_i32(np.frombuffer(pad, dtype=np.int32))


@lru_cache(maxsize=None)
def _load(group):
    return np.load(GOLDEN / f"{group}.npz", allow_pickle=False)


def names(group):
    return json.loads(str(_load(group)["__names__"]))


class Scenario:
    def __init__(self, group, name):
        z = _load(group)
        self.group, self.name = group, name
        g = lambda k: z[f"{name}|{k}"]
        self.meta = json.loads(str(g("meta")))
        self.source_text = str(g("source_text"))
        for k in ("actions", "state_hash", "obs_hash", "reward", "time", "flags", "n_possible", "samples",
                  "final_digest", "final_obs"):
            setattr(self, k, g(k))
        sc = g("li/scalars")
        specs = {}
        for sp in ("op_spec", "setup_spec", "travel_spec", "m_ofreq_spec", "m_odur_spec", "t_ofreq_spec", "t_odur_spec"):
            if f"{name}|li/{sp}/kind" in z.files:
                specs[sp] = L.TimeSpec(g(f"li/{sp}/kind"), g(f"li/{sp}/base"), g(f"li/{sp}/param"))
        fields = ("job_op_start", "op_machine", "op_tool", "op_duration", "pre_type", "pre_cap", "post_type",
                  "post_cap", "setup_time", "sbuf_type", "sbuf_cap", "sbuf_role", "buf_ref_id", "travel_time",
                  "agv_init_place", "job_init_buf", "m_outage_start", "m_outage_freq", "m_outage_dur",
                  "t_outage_freq", "t_outage_dur")
        self.lowered = L.LoweredInstance(
            description=name, n_jobs=int(sc[0]), n_machines=int(sc[1]), n_transports=int(sc[2]), n_sbuf=int(sc[3]),
            n_tools=int(sc[4]), default_tool=int(sc[5]), start_time=int(sc[6]), lower_bound=int(sc[7]),
            max_allowed_time=int(sc[8]), tools=json.loads(str(g("li/tools"))),
            **{k: g(f"li/{k}") for k in fields}, **specs)
        self.start_seeds = g("li/start_seeds") if f"{name}|li/start_seeds" in z.files else None

    @property
    def cfg(self) -> dict:
        c = dict(self.meta.get("cfg", {}))
        of = c.pop("observation_factory", None)
        if of == "BinaryOperationArrayObservation":
            c["obs_kind"] = 2
        elif of == "TasselJsspObservation":
            c["obs_kind"] = 3
        return c

    @property
    def cut(self) -> bool:
        """the harness stopped the episode at max_steps before it was done"""
        return self.meta.get("max_steps") is not None and self.meta["n_steps"] >= self.meta["max_steps"]

    @property
    def kind(self):
        return self.meta["kind"]

    @property
    def exception(self):
        return self.meta["exception"]


class KsFixture:
    """Reference makespan sample of a stochastic instance (oracle/gen_golden.py ks)."""

    def __init__(self, name):
        z = _load("ks")
        g = lambda k: z[f"{name}|{k}"]
        self.name = name
        self.makespan = g("makespan")
        self.n_steps = g("n_steps")
        self.source_text = str(g("source_text"))
        sc = g("li/scalars")
        specs = {}
        for sp in ("op_spec", "setup_spec", "travel_spec", "m_ofreq_spec", "m_odur_spec", "t_ofreq_spec", "t_odur_spec"):
            if f"{name}|li/{sp}/kind" in z.files:
                specs[sp] = L.TimeSpec(g(f"li/{sp}/kind"), g(f"li/{sp}/base"), g(f"li/{sp}/param"))
        fields = ("job_op_start", "op_machine", "op_tool", "op_duration", "pre_type", "pre_cap", "post_type",
                  "post_cap", "setup_time", "sbuf_type", "sbuf_cap", "sbuf_role", "buf_ref_id", "travel_time",
                  "agv_init_place", "job_init_buf", "m_outage_start", "m_outage_freq", "m_outage_dur",
                  "t_outage_freq", "t_outage_dur")
        self.lowered = L.LoweredInstance(
            description=name, n_jobs=int(sc[0]), n_machines=int(sc[1]), n_transports=int(sc[2]), n_sbuf=int(sc[3]),
            n_tools=int(sc[4]), default_tool=int(sc[5]), start_time=int(sc[6]), lower_bound=int(sc[7]),
            max_allowed_time=int(sc[8]), tools=json.loads(str(g("li/tools"))),
            **{k: g(f"li/{k}") for k in fields}, **specs)


def all_scenarios(groups=GROUPS):
    return [(g, n) for g in groups for n in names(g)]
