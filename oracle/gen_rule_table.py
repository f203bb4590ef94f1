"""Generate tests/golden/rules_all.npz: dispatch-rule rollouts of the UNMODIFIED reference over EVERY bundled
deterministic instance.

TEST INFRASTRUCTURE ONLY.  Usage (here, where /root/reference exists; ~25 min on 8 cores):

    python oracle/gen_rule_table.py [--procs 8] [--only ft06,la01]

For each of the spec files (data/jssp_instances/spec_files/*), transport YAMLs (data/jssp_instances/transport/*.yaml)
and deterministic DSL YAMLs (data/jssp_instances/dsl/*.yaml) the reference's own rule driver
(jobshoplab/utils/dipatching_benchmark.py:17-67, 131-137: `while not env.done: env.step(rule(env))`) is run with
the default Config() for fifo (always accept), spt and "mwkr"; the table stores per (instance, rule) the final
makespan (`env.state.state.time.time`), the number of env steps, the sum of rewards (f64) and the number of no-op
steps, plus the instance's source text so that the product compiler and the CUDA engine can be run on it where
/root/reference does not exist.  tests/test_rules_all.py checks the C oracle (CPU) and jsl_rollout (GPU,
throughput build) against every entry: "bit-exact makespans on all bundled deterministic instances".
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))

REF = Path("/root/reference")
DATA = REF / "data" / "jssp_instances"
OUT = HERE.parent / "tests" / "golden" / "rules_all.npz"
RULES = ("fifo", "spt", "mwkr")


def instance_list():
    out = []
    for p in sorted((DATA / "spec_files").iterdir()):
        if p.is_file():
            out.append(("spec", f"spec/{p.name}", p))
    for p in sorted((DATA / "transport").glob("*.yaml")):
        out.append(("dsl", f"transport/{p.stem}", p))
    for p in sorted((DATA / "dsl").glob("*.yaml")):
        out.append(("dsl", f"dsl/{p.stem}", p))
    return out


def run_one(task):
    kind, name, path, rule = task
    import ref_harness as rh

    t0 = time.time()
    try:
        env = rh.make_env(kind, path)
    except Exception as ex:  # an instance the reference itself cannot compile
        return name, rule, {"skip": f"compile: {type(ex).__name__}"}
    li = env.jsl_rc.lowered
    if any(getattr(li, sp).kind is not None for sp in ("op_spec", "setup_spec", "travel_spec", "m_ofreq_spec",
                                                        "m_odur_spec", "t_ofreq_spec", "t_odur_spec")):
        return name, rule, {"skip": "stochastic"}
    ret, steps, noops, exc = 0.0, 0, 0, ""
    while not env.done:
        a = rh.rule_action(env, rule)
        try:
            _, r, _, _, _ = env.step(a)
        except Exception as ex:
            exc = type(ex).__name__
            break
        ret += r
        steps += 1
        noops += a == 0
    return name, rule, {"makespan": int(env.state.state.time.time), "n_steps": steps, "ret": ret, "noops": int(noops),
                        "exception": exc, "terminated": bool(env.terminated), "truncated": bool(env.truncated),
                        "shape": [li.n_jobs, li.n_machines, li.n_transports, li.n_sbuf],
                        "lb_mat": [li.lower_bound, li.max_allowed_time], "secs": time.time() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    insts = instance_list()
    if args.only:
        keep = set(args.only.split(","))
        insts = [i for i in insts if i[1].split("/")[1] in keep]
    # biggest first so the pool drains evenly
    insts.sort(key=lambda i: -i[2].stat().st_size)
    tasks = [(k, n, str(p), r) for k, n, p in insts for r in RULES]
    res = {}
    t0 = time.time()
    with mp.Pool(args.procs) as pool:
        for i, (name, rule, r) in enumerate(pool.imap_unordered(run_one, tasks)):
            res[(name, rule)] = r
            print(f"[{i + 1}/{len(tasks)}] {name} {rule}: {r} ({time.time() - t0:.0f}s)", flush=True)
    names, kinds, texts, table, rets, meta = [], [], [], [], [], {}
    for kind, name, path in sorted(insts, key=lambda i: i[1]):
        rs = [res[(name, r)] for r in RULES]
        if any("skip" in r for r in rs):
            meta.setdefault("skipped", {})[name] = rs[0].get("skip")
            continue
        names.append(name); kinds.append(kind); texts.append(Path(path).read_text())
        table.append([[r["makespan"], r["n_steps"], r["noops"], int(r["terminated"]), int(r["truncated"])] for r in rs])
        rets.append([r["ret"] for r in rs])
        meta[name] = {"shape": rs[0]["shape"], "lb_mat": rs[0]["lb_mat"], "exception": [r["exception"] for r in rs]}
    np.savez_compressed(OUT, names=np.asarray(json.dumps(names)), kinds=np.asarray(json.dumps(kinds)),
                        texts=np.asarray(json.dumps(texts)), rules=np.asarray(json.dumps(RULES)),
                        table=np.asarray(table, dtype=np.int64), ret=np.asarray(rets, dtype=np.float64),
                        meta=np.asarray(json.dumps(meta)))
    print(f"wrote {OUT}: {len(names)} instances x {len(RULES)} rules in {time.time() - t0:.0f}s; "
          f"skipped {meta.get('skipped', {})}")


if __name__ == "__main__":
    main()
