"""Differential fuzzing of the oracle (and the product compiler) against the UNMODIFIED reference on random
DSL instances -- TEST INFRASTRUCTURE, runs only where /root/reference exists.

    python oracle/fuzz_reference.py [n] [seed]

Random YAML instances: shapes up to 6x5, random travel matrices (with zeros), FIFO/LIFO/DUMMY/FLEX machine
buffers with random capacities (global or per machine), tool usage + static setup matrices, deterministic
outages, explicit AGV fleets with init locations, early transport on/off.  Every step compares the canonical
digest, reward, flags and observation; exceptions must map to the same status.
"""
import sys
from pathlib import Path

import numpy as np
import yaml

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))
import ref_harness as rh  # noqa: E402
import oracle_py as op  # noqa: E402
from jobshoplab_b200 import compiler as C  # noqa: E402

EXC = {"BufferFullError": 4, "UnsuccessfulStateMachineResult": 5, "InvalidValue": 6, "NotImplementedError": 9,
       "TravelTimeError": 9, "TransportConfigError": 9, "MissingProcessingOperationError": 9, "JobNotInBufferError": 9}


def random_dsl(rng) -> dict:
    J, M = int(rng.integers(2, 7)), int(rng.integers(2, 6))
    lines = ["|".join(f"(m{m},t)" for m in range(M))]
    for j in range(J):
        perm = rng.permutation(M)
        lines.append(f"j{j}|" + " ".join(f"({int(m)},{int(rng.integers(1, 10))})" for m in perm))
    ic = {"description": "fuzz", "instance": {"description": "fuzz", "specification": "\n".join(lines) + "\n"}}
    d = {"title": "InstanceConfig", "instance_config": ic}
    names = [f"m-{m}" for m in range(M)] + ["in-buf", "out-buf"]
    if rng.random() < 0.8:
        mat = rng.integers(0, 7, size=(M + 2, M + 2))
        mat[rng.random(mat.shape) < 0.25] = 0
        np.fill_diagonal(mat, 0)
        spec = "|".join(names) + "\n" + "\n".join(f"{names[i]}|" + " ".join(str(int(v)) for v in mat[i]) for i in range(M + 2)) + "\n"
        lg = {"specification": spec}
        if rng.random() < 0.5:
            lg["type"] = "agv"; lg["amount"] = int(rng.integers(1, J + 2))
        ic["logistics"] = lg
    types = ["fifo", "lifo", "dummy", "flex"]
    if rng.random() < 0.6:
        def b():
            return [{"capacity": int(rng.choice([1, 2, 3, J, J + 2])), "type": str(rng.choice(types))}]
        if rng.random() < 0.6:
            ic["machines"] = {"prebuffer": b(), "postbuffer": b()}
        else:
            ic["machines"] = [{f"m-{m}": {"prebuffer": b(), "postbuffer": b()}} for m in range(M) if rng.random() < 0.7]
    if rng.random() < 0.4:
        nt = int(rng.integers(2, 4))
        tools = [f"tl-{i}" for i in range(nt)]
        ic["instance"]["tool_usage"] = [{"job": f"j{j}", "operation_tools": [str(rng.choice(tools)) for _ in range(M)]}
                                        for j in range(J)]
        ic["setup_times"] = []
        for m in range(M):
            mat = rng.integers(0, 5, size=(nt, nt))
            spec = "|".join(tools) + "\n" + "\n".join(f"{tools[i]}|" + " ".join(str(int(v)) for v in mat[i]) for i in range(nt)) + "\n"
            ic["setup_times"].append({"machine": f"m-{m}", "specification": spec, "time_behavior": "static"})
    if rng.random() < 0.4:
        outs = []
        if rng.random() < 0.8:
            outs.append({"component": "m", "type": "maintenance", "duration": int(rng.integers(0, 6)), "frequency": int(rng.integers(0, 40))})
        if rng.random() < 0.3:
            outs.append({"component": f"m-{int(rng.integers(0, M))}", "type": "fail", "duration": int(rng.integers(1, 4)), "frequency": int(rng.integers(5, 30))})
        if rng.random() < 0.5:
            outs.append({"component": "t", "type": "recharge", "duration": int(rng.integers(0, 5)), "frequency": int(rng.integers(0, 30))})
        ic["outages"] = outs
    if rng.random() < 0.3:
        init = {}
        T = ic.get("logistics", {}).get("amount", J)
        for t in range(T):
            if rng.random() < 0.5:
                init[f"t-{t}"] = {"location": f"m-{int(rng.integers(0, M))}"}
        if init:
            d["init_state"] = init
    return d


def run_one(rng, max_steps=500):
    d = random_dsl(rng)
    cfg = dict(allow_early_transport=bool(rng.random() < 0.6))
    text = yaml.safe_dump(d, sort_keys=False)
    try:
        env = rh.make_env("dslstr", text, **cfg)
    except Exception as ex:  # the reference cannot even reset (e.g. buffer overflow at reset)
        return 0, f"ctor:{type(ex).__name__}", text
    ix = rh.Indexer(env.instance)
    li = env.jsl_rc.lowered
    mine = C.compile_dsl_text(text)
    for k in ("op_machine", "op_duration", "op_tool", "pre_type", "pre_cap", "post_type", "post_cap", "setup_time",
              "travel_time", "agv_init_place", "job_init_buf", "m_outage_start", "m_outage_freq", "m_outage_dur",
              "t_outage_freq", "t_outage_dur", "buf_ref_id"):
        assert np.array_equal(getattr(mine, k), getattr(li, k)), ("compiler", k)
    o = op.OracleEnv(mine, **cfg)
    assert np.array_equal(rh.digest(ix, env.state), o.digest()), "reset"
    p = float(rng.choice([0.5, 0.8, 1.0]))
    n = 0
    while not env.done and n < max_steps:
        a = int(rng.random() < p)
        try:
            obs, r, term, trunc, info = env.step(a)
        except Exception as ex:
            st = o.step(a)
            name = type(ex).__name__
            assert st == EXC.get(name, 9), (name, st)
            return n, name, text
        st = o.step(a)
        assert st <= 3, st
        dr, do = rh.digest(ix, env.state), o.digest()
        assert np.array_equal(dr, do), f"step {n} a={a}\n ref {dr.tolist()}\n ora {do.tolist()}"
        assert r == o.reward and term == o.terminated and trunc == o.truncated, (n, r, o.reward)
        assert rh.obs_bytes(obs) == o.obs(), (n, "obs")
        n += 1
    return n, "done" if env.done else "cut", text


def main(n=200, seed=0):
    rng = np.random.default_rng(seed)
    total, hist = 0, {}
    for i in range(n):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        try:
            steps, outcome, text = run_one(sub)
        except AssertionError as ex:
            print(f"FUZZ FAILURE at instance {i} (seed {seed}): {str(ex)[:3000]}")
            raise SystemExit(1)
        total += steps
        hist[outcome] = hist.get(outcome, 0) + 1
    print(f"reference fuzz ok: {n} instances, {total} steps, outcomes {dict(sorted(hist.items()))}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
