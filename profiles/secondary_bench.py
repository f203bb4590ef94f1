"""Throughput + correctness of the other BASELINE.json configs on one GPU (rollout mode, 65 536 replicas):
ft10 / la16 SPT / MWKR (bit-exact makespans), transport ft10-t / la01-t with random actions, the stochastic
ft20-t variant (setup times, breakdowns, Poisson travel), FIFO buffers with capacity 5 (BASELINE config 4).
Prints one JSON line per workload.

    python profiles/secondary_bench.py [B]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _golden as G  # noqa: E402
from jobshoplab_b200 import compiler, engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
EXPECT = {("ft10", "spt"): 1074, ("ft10", "mwkr"): 1262, ("ft10", "fifo"): 1262, ("la16", "spt"): 1156,
          ("la16", "mwkr"): 1268, ("la16", "fifo"): 1230, ("ta01", "spt"): 1462, ("ta41", "spt"): 2465,
          ("ta41", "mwkr"): 3120, ("ta41", "fifo"): 2973}


def timed(batch, policy, reps=3):
    for _ in range(2):
        r = batch.rollout(policy, reset_first=True, writeback=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = batch.rollout(policy, reset_first=True, writeback=False)
    e1.record()
    torch.cuda.synchronize()
    steps = int(r.n_steps.sum())
    return r, steps * reps / (e0.elapsed_time(e1) / 1e3)


def main():
    import os
    only_general = os.environ.get("JSL_SECONDARY") in ("general", "stoch")
    only_stoch = os.environ.get("JSL_SECONDARY") == "stoch"
    for inst in (() if only_general else ("ft10", "la16", "ta01", "ta41")):
        li = compiler.compile_spec_text(G.Scenario("classic", f"{inst}/fifo").source_text)
        batch = engine.Batch(li, B)
        for pol in ("spt", "mwkr", "fifo", "random"):
            r, rate = timed(batch, pol)
            mk = r.makespan
            ok = None
            if (inst, pol) in EXPECT:
                ok = bool((mk == EXPECT[(inst, pol)]).all())
            print(json.dumps({"workload": f"{inst} x{B} {pol}", "env_steps_per_sec": rate, "makespan_min": int(mk.min()),
                              "makespan_max": int(mk.max()), "bit_exact_vs_reference": ok,
                              "terminated": int((r.status == 1).sum())}))
        batch.close()
    for name in (() if only_stoch else ("ft10-t", "la01-t", "ta41-t")):
        li = compiler.compile_dsl_text(G.Scenario("transport", f"{name}/fifo").source_text)
        batch = engine.Batch(li, B)
        for pol in ("fifo", "random"):
            r, rate = timed(batch, pol)
            exp = G.Scenario("transport", f"{name}/fifo").meta["makespan"] if pol == "fifo" else None
            print(json.dumps({"workload": f"{name} x{B} {pol}", "env_steps_per_sec": rate, "makespan_min": int(r.makespan.min()),
                              "makespan_max": int(r.makespan.max()),
                              "bit_exact_vs_reference": None if exp is None else bool((r.makespan == exp).all()),
                              "terminated": int((r.status == 1).sum())}))
        batch.close()
    # BASELINE config 4: FIFO pre/post buffers with capacity 5 (general kernel); random actions may run a buffer full
    for name in (() if only_stoch else ("la01-t/fifo5/random0", "ft06-t/fifo5/random0")):
        sc = G.Scenario("buffers", name)
        batch = engine.Batch(sc.lowered, B, seed=sc.meta["seed"], **sc.cfg)
        r, rate = timed(batch, "random")
        st = r.status.cpu().numpy()
        e0 = sc.meta["env_idx"]
        print(json.dumps({"workload": f"{name.rsplit('/', 1)[0]} (FIFO buffers, capacity 5) x{B} random",
                          "env_steps_per_sec": rate, "terminated": int((st == 1).sum()),
                          "buffer_full": int((st == 4).sum()), "other_status": int(((st != 1) & (st != 4)).sum()),
                          "bit_exact_vs_reference": bool(int(r.makespan[e0]) == sc.meta["makespan"] and
                                                         int(r.n_steps[e0]) == sc.meta["n_steps"])}))
        batch.close()
    fx = G.KsFixture("ft20-t_features")
    batch = engine.Batch(fx.lowered, B, seed=5)
    r, rate = timed(batch, "fifo")
    print(json.dumps({"workload": f"ft20-t setup+breakdowns+poisson travel x{B} fifo (stochastic, table mode)",
                      "env_steps_per_sec": rate, "makespan_mean": float(r.makespan.double().mean()),
                      "reference_makespan_mean": float(fx.makespan.mean()), "terminated": int((r.status == 1).sum())}))


if __name__ == "__main__":
    main()
