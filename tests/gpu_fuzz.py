"""Differential fuzzing of the CUDA engine ON THE GPU (through the C ABI) against the CPU oracle: random instances
of every specialisation level (tests/fuzz_engine.random_instance), a few envs per instance with random action tapes,
rollout mode; final status, step count, makespan, return and canonical digest must agree.  The CPU debug build
(tests/hostsim) runs the same engine source with loop primitives, so only this run sees real warp execution of the
lane-parallel code.

    python tests/gpu_fuzz.py [n_instances] [seed]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fuzz_engine  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def check_instance(eng, torch, li, cfg, sub, B=6, L=600, tag=""):
    p = float(sub.choice([0.5, 0.75, 1.0]))
    tape = (sub.random((B, L)) < p).astype(np.uint8)
    batch = eng.Batch(li, B, record_ops=True, **cfg)
    batch.reset()
    r = batch.rollout("tape", tape=torch.from_numpy(tape).cuda())
    mk, ns = r.makespan.cpu().numpy(), r.n_steps.cpu().numpy()
    ret, st = r.ret.cpu().numpy(), r.status.cpu().numpy()
    steps = 0
    for e in range(B):
        o = op.OracleEnv(li, **cfg)
        n = 0
        while n < L and o.status == 0:
            o.step(int(tape[e, n]))
            n += 1
        t = f"{tag} (J={li.n_jobs}, M={li.n_machines}, T={li.n_transports}) env {e}"
        assert st[e] == o.status, (t, "status", int(st[e]), o.status)
        assert ns[e] == o.n_steps, (t, "steps", int(ns[e]), o.n_steps)
        steps += o.n_steps
        if o.status > 3:
            continue  # an escaping exception: the reference has no state to compare
        assert mk[e] == o.time, (t, "time", int(mk[e]), o.time)
        assert ret[e] == o.ret, (t, "return")
        assert np.array_equal(batch.get_state(e), o.digest()), (t, "digest")
    batch.close()
    return steps


def main(n=200, seed=0, jrange=(1, 8)):
    import torch
    from jobshoplab_b200 import engine as eng

    rng = np.random.default_rng(seed)
    total = 0
    for i in range(n):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        li, cfg = fuzz_engine.random_instance(sub, simple=(None, 3, 1, 2, 4)[i % 5], jrange=jrange)
        total += check_instance(eng, torch, li, cfg, sub, tag=f"gpu fuzz instance {i} seed {seed}")
    print(f"gpu fuzz ok: {n} instances, {total} env-steps compared")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
