"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Usage (here, where /root/reference exists):

    python oracle/gen_golden.py [group ...]        # groups: classic solutions transport buffers
                                                   #         features truncation oparray

Each scenario stores: the reference-lowered instance arrays, the source text, the action tape, and
per step the 64-bit hash of the canonical state digest (oracle/ref_harness.py), the hash of the packed
observation, reward (f64), time, flags and candidate count; plus the stochastic sample tape, the final
digest and the name of an escaping exception, if any.  tests/test_oracle_golden.py replays the tapes
through the C oracle, tests/test_cuda_parity.py through the CUDA engine.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np
import yaml

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))

REF = Path("/root/reference")
DATA = REF / "data" / "jssp_instances"
TDATA = REF / "tests" / "data" / "jssp_instances"
OUT = HERE.parent / "tests" / "golden"

MASK = (1 << 64) - 1
P = 1099511628211
H0 = 1469598103934665603


def hash_i32(a: np.ndarray) -> int:
    """h = H0; for x: h = h*P + (u32)x + 1  (mod 2^64).  Same in tests/_hash.py and the C side."""
    h = H0
    for x in np.asarray(a, dtype=np.int32).view(np.uint32).tolist():
        h = (h * P + x + 1) & MASK
    return h


def hash_bytes(b: bytes) -> int:
    pad = b + b"\0" * (-len(b) % 4)
    return hash_i32(np.frombuffer(pad, dtype=np.int32))


LOWERED_FIELDS = ("job_op_start", "op_machine", "op_tool", "op_duration", "pre_type", "pre_cap", "post_type",
                  "post_cap", "setup_time", "sbuf_type", "sbuf_cap", "sbuf_role", "buf_ref_id", "travel_time",
                  "agv_init_place", "job_init_buf", "m_outage_start", "m_outage_freq", "m_outage_dur",
                  "t_outage_freq", "t_outage_dur")
SPEC_NAMES = ("op_spec", "setup_spec", "travel_spec", "m_ofreq_spec", "m_odur_spec", "t_ofreq_spec", "t_odur_spec")


def pack_lowered(li) -> dict:
    out = {f"li/{k}": getattr(li, k) for k in LOWERED_FIELDS}
    out["li/scalars"] = np.asarray([li.n_jobs, li.n_machines, li.n_transports, li.n_sbuf, li.n_tools,
                                    li.default_tool, li.start_time, li.lower_bound, li.max_allowed_time],
                                   dtype=np.int64)
    out["li/tools"] = np.asarray(json.dumps(li.tools))
    if getattr(li, "start_seeds", None) is not None:
        out["li/start_seeds"] = li.start_seeds
    for name in SPEC_NAMES:
        sp = getattr(li, name)
        if sp.kind is not None:
            out[f"li/{name}/kind"], out[f"li/{name}/base"], out[f"li/{name}/param"] = sp.kind, sp.base, sp.param
    return out


def run_scenario(sc: dict) -> dict:
    import ref_harness as rh

    kind, src = sc["kind"], sc["src"]
    cfg = sc.get("cfg", {})
    if kind == "spec":
        text = Path(src).read_text()
        env = rh.make_env("spec", src, **cfg)
    else:
        d = yaml.safe_load(Path(src).read_text()) if isinstance(src, (str, Path)) else src
        for path, val in sc.get("patch", []):
            cur = d
            for k in path[:-1]:
                cur = cur.setdefault(k, {})
            if val is None:
                cur.pop(path[-1], None)
            else:
                cur[path[-1]] = val
        text = yaml.safe_dump(d, sort_keys=False)
        env = rh.make_env("dslstr", text, **cfg)
    ix = rh.Indexer(env.instance)
    li = env.jsl_rc.lowered
    of = cfg.get("observation_factory")
    obs_fn = {"BinaryOperationArrayObservation": rh.obs_oparray_bytes,
              "TasselJsspObservation": rh.obs_tassel_bytes}.get(of, rh.obs_bytes)
    policy = sc["policy"]
    state_hash = [hash_i32(rh.digest(ix, env.state))]
    obs_hash = [hash_bytes(obs_fn(env.current_observation[0]))]
    times = [env.state.state.time.time]
    npos = [len(env.state.possible_transitions)]
    actions, rewards, flags = [], [], []
    exc = ""
    sol_iter = None
    if policy == "solution":
        from jobshoplab.utils import solutions
        seq = sorted(tuple(solutions.make_solution_action_sequence(sc["solution"])), key=lambda x: (x[2], x[1], x[0]))
        expected_makespan = solutions.get_make_span(seq, env.instance)
        sol_iter = iter(seq)
        cur = next(sol_iter)
    step = 0
    last_obs = env.current_observation[0]
    while not env.done and step < sc.get("max_steps", 1 << 30):
        if policy == "solution":
            # protocol of tests/end_to_end_tests/test_env_end_to_end.py:72-105
            machine, job, start = cur
            if env.state.state.time.time < start:
                a = 0
            else:
                ev = env.state.possible_transitions[0]
                if int(ev.job_id.split("-")[1]) == job:
                    assert env.state.state.time.time == start
                    assert int(ev.component_id.split("-")[1]) == machine
                    a = 1
                    cur = next(sol_iter, None)
                else:
                    a = 0
        elif policy == "random":
            a = rh.random_action(sc["seed"], sc.get("env_idx", 0), step)
        elif policy == "tape":
            a = sc["tape"][step]
        else:
            a = rh.rule_action(env, policy)
        actions.append(a)
        try:
            obs, r, term, trunc, info = env.step(a)
        except Exception as ex:  # escapes env.step in the reference
            exc = type(ex).__name__
            break
        last_obs = obs
        step += 1
        rewards.append(r)
        flags.append(int(term) | (int(trunc) << 1))
        state_hash.append(hash_i32(rh.digest(ix, env.state)))
        obs_hash.append(hash_bytes(obs_fn(obs)))
        times.append(env.state.state.time.time)
        npos.append(len(env.state.possible_transitions))
    if policy == "solution" and not exc:
        assert env.state.state.time.time == expected_makespan, (sc["name"], env.state.state.time.time, expected_makespan)
        assert env.terminated
    rec = pack_lowered(li)
    meta = {k: v for k, v in sc.items() if k not in ("tape", "src")}
    meta["src"] = str(src) if isinstance(src, (str, Path)) else "<dict>"
    meta["exception"] = exc
    meta["makespan"] = int(env.state.state.time.time)
    meta["n_steps"] = step
    meta["no_op_count"] = int(sum(int(len(h.action.transitions) == 0) for h in env.history))
    rec.update({
        "meta": np.asarray(json.dumps(meta)),
        "source_text": np.asarray(text),
        "actions": np.asarray(actions, dtype=np.uint8),
        "state_hash": np.asarray(state_hash, dtype=np.uint64),
        "obs_hash": np.asarray(obs_hash, dtype=np.uint64),
        "reward": np.asarray(rewards, dtype=np.float64),
        "time": np.asarray(times, dtype=np.int32),
        "flags": np.asarray(flags, dtype=np.uint8),
        "n_possible": np.asarray(npos, dtype=np.int32),
        "samples": np.asarray(list(rh.SAMPLE_TAPE), dtype=np.int32),
        "final_digest": rh.digest(ix, env.state),
        "final_obs": np.frombuffer(obs_fn(last_obs), dtype=np.uint8),
    })
    return {"name": sc["name"], "rec": rec}


# ------------------------------------------------------------------ scenario lists
def spec(name):
    return str(DATA / "spec_files" / name)


def g_classic():
    out = []
    for name in ["2x2", "3x3", "ft06", "ft10", "la16", "ft20", "abz7", "la21", "swv16", "ta01", "ta41", "yn1"]:
        for pol in ["fifo", "spt", "mwkr"]:
            out.append(dict(name=f"{name}/{pol}", kind="spec", src=spec(name), policy=pol))
        for seed in ([0, 1, 2] if name not in ("ta41", "yn1", "swv16") else [0]):
            out.append(dict(name=f"{name}/random{seed}", kind="spec", src=spec(name), policy="random", seed=seed,
                            env_idx=7 * seed))
    return out


def g_solutions():
    out = []
    for f in sorted(os.listdir(DATA / "spec_files" / "solutions")):
        inst = f.split("_")[0]
        out.append(dict(name=f"{inst}/solution", kind="spec", src=spec(inst), policy="solution",
                        solution=str(DATA / "spec_files" / "solutions" / f)))
    return out


def g_transport():
    out = []
    for name in ["ft06", "ft10", "la01", "la02", "ft20", "la16", "la21", "ta41"]:
        for early in (True, False):
            src = str(DATA / "transport" / f"{name}.yaml")
            tag = "" if early else "/noearly"
            for pol in ["fifo", "spt", "mwkr"]:
                out.append(dict(name=f"{name}-t{tag}/{pol}", kind="dsl", src=src, policy=pol,
                                cfg=dict(allow_early_transport=early)))
            for seed in [0, 1]:
                out.append(dict(name=f"{name}-t{tag}/random{seed}", kind="dsl", src=src, policy="random", seed=seed,
                                env_idx=3 + seed, cfg=dict(allow_early_transport=early)))
    return out


def g_buffers():
    out = []
    for name in ["la01", "ft06"]:
        src = str(DATA / "transport" / f"{name}.yaml")
        for typ in ["fifo", "lifo", "dummy", "flex"]:
            for c in [1, 2, 3, 5]:
                for early in (True, False):
                    patch = [(["instance_config", "machines"],
                              {"prebuffer": [{"capacity": c, "type": typ}], "postbuffer": [{"capacity": c, "type": typ}]})]
                    tag = f"{name}-t/{typ}{c}{'' if early else '/noearly'}"
                    for seed in [0, 1, 2]:
                        out.append(dict(name=f"{tag}/random{seed}", kind="dsl", src=src, patch=patch, policy="random",
                                        seed=seed, env_idx=seed, cfg=dict(allow_early_transport=early)))
                    out.append(dict(name=f"{tag}/fifo", kind="dsl", src=src, patch=patch, policy="fifo",
                                    cfg=dict(allow_early_transport=early)))
    return out


def g_features():
    out = []
    paths = [TDATA / "full_feature_3x3_instance.yaml", DATA / "dsl" / "instance_3x3_full.yaml",
             TDATA / "instance_with_agvs.yaml", DATA / "dsl" / "real_world_deterministic_6x6_instance.yaml",
             DATA / "dsl" / "real_world_stochastic_6x6_instance.yaml", DATA / "dsl" / "real_world_instance.yaml",
             DATA / "dsl" / "default_instance.yaml", DATA / "dsl" / "minimal_instance.yaml",
             DATA / "dsl" / "minimal_instance_with_transport.yaml"]
    for p in paths:
        for early in (True, False):
            tag = f"{p.stem}{'' if early else '/noearly'}"
            for pol in ["fifo", "spt", "mwkr"]:
                out.append(dict(name=f"{tag}/{pol}", kind="dsl", src=str(p), policy=pol, cfg=dict(allow_early_transport=early)))
            for seed in [0, 1, 2]:
                out.append(dict(name=f"{tag}/random{seed}", kind="dsl", src=str(p), policy="random", seed=seed,
                                env_idx=seed, cfg=dict(allow_early_transport=early)))
    # deterministic setup times + deterministic outages (strip every stochastic behaviour)
    det_patch = [(["instance_config", "logistics", "time_behavior"], None),
                 (["instance_config", "outages"], [
                     {"component": "m", "type": "maintenance", "duration": 5, "frequency": 20},
                     {"component": "t", "type": "recharge", "duration": 3, "frequency": 10}])]
    for p in [TDATA / "full_feature_3x3_instance.yaml"]:
        d = yaml.safe_load(p.read_text())
        for st in d["instance_config"]["setup_times"]:
            st["time_behavior"] = "static"
        for pol, seed in [("fifo", 0), ("spt", 0), ("random", 0), ("random", 1), ("random", 2)]:
            out.append(dict(name=f"{p.stem}/det/{pol}{seed}", kind="dsl", src=d, patch=det_patch, policy=pol, seed=seed,
                            env_idx=seed))
    return out


def g_truncation():
    out = []
    for name in ["ft06", "la16"]:
        for joker in (0, 2, 5):
            for seed in range(4):
                out.append(dict(name=f"{name}/trunc{joker}/random{seed}", kind="spec", src=spec(name), policy="random",
                                seed=seed, env_idx=seed, cfg=dict(truncation_active=True, truncation_joker=joker)))
    src = str(DATA / "transport" / "ft06.yaml")
    for seed in range(4):
        out.append(dict(name=f"ft06-t/trunc1/random{seed}", kind="dsl", src=src, policy="random", seed=seed,
                        env_idx=seed, cfg=dict(truncation_active=True, truncation_joker=1)))
    for seed in range(2):
        out.append(dict(name=f"ft06/biases/random{seed}", kind="spec", src=spec("ft06"), policy="random", seed=seed,
                        env_idx=seed, cfg=dict(truncation_active=True, truncation_joker=0, sparse_bias=2.0,
                                               dense_bias=0.5, truncation_bias=-3.0)))
    return out


def g_oparray():
    out = []
    cfg = dict(observation_factory="BinaryOperationArrayObservation")
    for name in ["ft06", "ta01"]:
        for pol, seed in [("fifo", 0), ("random", 0)]:
            out.append(dict(name=f"{name}/oparray/{pol}{seed}", kind="spec", src=spec(name), policy=pol, seed=seed, cfg=cfg))
    src = str(DATA / "transport" / "ft10.yaml")
    for pol, seed in [("fifo", 0), ("random", 0)]:
        out.append(dict(name=f"ft10-t/oparray/{pol}{seed}", kind="dsl", src=src, policy=pol, seed=seed, cfg=cfg))
    return out


def g_tassel():
    out = []
    cfg = dict(observation_factory="TasselJsspObservation")
    for name in ["ft06", "la16", "ta01"]:
        for pol, seed in [("fifo", 0), ("spt", 0), ("random", 0), ("random", 1)]:
            out.append(dict(name=f"{name}/tassel/{pol}{seed}", kind="spec", src=spec(name), policy=pol, seed=seed, cfg=cfg))
    for name in ["ft10", "la01"]:
        src = str(DATA / "transport" / f"{name}.yaml")
        for pol, seed in [("fifo", 0), ("random", 0)]:
            out.append(dict(name=f"{name}-t/tassel/{pol}{seed}", kind="dsl", src=src, policy=pol, seed=seed, cfg=cfg))
    p = TDATA / "full_feature_3x3_instance.yaml"
    for seed in range(3):
        out.append(dict(name=f"{p.stem}/tassel/random{seed}", kind="dsl", src=str(p), policy="random", seed=seed, cfg=cfg))
    return out


def g_big():
    """100 jobs x 20 machines (ta71): exercises the 4-lane-word kernels (JW = 4)."""
    return [dict(name="ta71/fifo", kind="spec", src=spec("ta71"), policy="fifo"),
            dict(name="ta71/random0", kind="spec", src=spec("ta71"), policy="random", seed=0, env_idx=5, max_steps=700)]


GROUPS = {"big": g_big, "tassel": g_tassel, "classic": g_classic, "solutions": g_solutions, "transport": g_transport, "buffers": g_buffers,
          "features": g_features, "truncation": g_truncation, "oparray": g_oparray}


def _safe(sc):
    try:
        return run_scenario(sc)
    except Exception as ex:  # keep the pool alive; report
        import traceback
        return {"name": sc["name"], "error": f"{type(ex).__name__}: {ex}\n{traceback.format_exc()}"}


def main(argv):
    groups = [g for g in argv if g != "ks"] if argv else list(GROUPS)
    if argv and not groups:
        return
    OUT.mkdir(parents=True, exist_ok=True)
    for g in groups:
        scs = GROUPS[g]()
        t0 = time.time()
        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            results = pool.map(_safe, scs, chunksize=1)
        flat = {}
        names = []
        for r in results:
            if "error" in r:
                print("ERROR", r["name"], r["error"])
                continue
            names.append(r["name"])
            for k, v in r["rec"].items():
                flat[f"{r['name']}|{k}"] = v
        flat["__names__"] = np.asarray(json.dumps(names))
        np.savez_compressed(OUT / f"{g}.npz", **flat)
        steps = sum(len(flat[f"{n}|actions"]) for n in names)
        print(f"{g}: {len(names)}/{len(scs)} scenarios, {steps} steps, {time.time() - t0:.0f}s, "
              f"{(OUT / f'{g}.npz').stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main(sys.argv[1:])


# ------------------------------------------------------------------ distribution fixtures (KS tests)
def _ks_episode(args):
    import ref_harness as rh

    text, seed, policy = args
    env = rh.make_env("dslstr", text, seed=seed)
    n = 0
    while not env.done:
        a = 1 if policy == "fifo" else rh.random_action(seed, 0, n)
        env.step(a)
        n += 1
    return int(env.state.state.time.time), n, int(env.terminated)


def ks_sources():
    full = (TDATA / "full_feature_3x3_instance.yaml").read_text()
    d = yaml.safe_load((DATA / "transport" / "ft20.yaml").read_text())
    ic = d["instance_config"]
    ic["logistics"]["time_behavior"] = {"type": "poisson"}
    ic["outages"] = [{"component": "m", "type": "breakdown", "duration": 5,
                      "frequency": {"type": "gamma", "scale": 5, "base": 10}}]
    tools = ["tl-0", "tl-1", "tl-2"]
    ic["instance"]["tool_usage"] = [{"job": f"j{j}", "operation_tools": [tools[(j + k) % 3] for k in range(5)]}
                                    for j in range(20)]
    ic["setup_times"] = [{"machine": f"m-{m}", "specification": "tl-0|tl-1|tl-2\ntl-0|0 2 5\ntl-1|2 0 8\ntl-2|5 2 0\n",
                          "time_behavior": {"type": "uni", "offset": 2}} for m in range(5)]
    return {"full_feature_3x3": (full, 3000), "ft20-t_features": (yaml.safe_dump(d, sort_keys=False), 1200)}


def gen_ks():
    import ref_harness as rh

    out = {}
    for name, (text, n) in ks_sources().items():
        t0 = time.time()
        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            res = pool.map(_ks_episode, [(text, s, "fifo") for s in range(n)], chunksize=8)
        mk = np.asarray([r[0] for r in res], dtype=np.int32)
        env = rh.make_env("dslstr", text, seed=0)
        for k, v in pack_lowered(env.jsl_rc.lowered).items():
            out[f"{name}|{k}"] = v
        out[f"{name}|makespan"] = np.sort(mk)
        out[f"{name}|n_steps"] = np.asarray([r[1] for r in res], dtype=np.int32)
        out[f"{name}|terminated"] = np.asarray([r[2] for r in res], dtype=np.uint8)
        out[f"{name}|source_text"] = np.asarray(text)
        print(f"ks/{name}: {n} reference episodes, makespan mean {mk.mean():.1f} std {mk.std():.1f}, {time.time() - t0:.0f}s")
    out["__names__"] = np.asarray(json.dumps(list(ks_sources())))
    np.savez_compressed(OUT / "ks.npz", **out)


if __name__ == "__main__" and "ks" in sys.argv[1:]:
    gen_ks()
