"""Differential fuzzing of the CUDA engine ON THE GPU (through the C ABI) against the CPU oracle: random instances
of every specialisation level (tests/fuzz_engine.random_instance), a few envs per instance with random action tapes.
The CPU debug build (tests/hostsim) runs the same engine source with loop primitives, so only this run sees real warp
execution of the lane-parallel code.

Both builds of the library are pinned:
  * libjsl_b200_log.so (record_ops batches): rollout mode -- final status, step count, makespan, return, canonical
    digest;
  * libjsl_b200.so, the throughput build bench.py measures: rollout mode -- the same, with the digest's DONE-operation
    times masked (that build keeps no per-op log) -- and step mode -- status, reward, time, flags, candidate and the
    packed observation of EVERY env.step.

    python tests/gpu_fuzz.py [n_instances] [seed]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fuzz_engine  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def mask_done_ops(digest: np.ndarray, li) -> np.ndarray:
    """Canonical digest with the (start, end) of DONE operations set to -1: what a build without the per-op log
    can know (layout: DESIGN.md 'canonical digest')."""
    d = digest.copy()
    pos = 5
    for j in range(li.n_jobs):
        pos += 3
        for _ in range(int(li.job_op_start[j + 1]) - int(li.job_op_start[j])):
            if d[pos] == 2:
                d[pos + 1] = d[pos + 2] = -1
            pos += 3
    return d


def make_tape(sub, B, L):
    p = float(sub.choice([0.5, 0.75, 1.0]))
    return (sub.random((B, L)) < p).astype(np.uint8)


def check_instance(eng, torch, li, cfg, sub, B=6, L=600, tag="", record=True, tape=None):
    """Rollout mode against the oracle.  record=True: logging build, full digest; False: throughput build."""
    tape = make_tape(sub, B, L) if tape is None else tape
    batch = eng.Batch(li, B, record_ops=record, **cfg)
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
        t = f"{tag} (J={li.n_jobs}, M={li.n_machines}, T={li.n_transports}, log build={record}) env {e}"
        assert st[e] == o.status, (t, "status", int(st[e]), o.status)
        assert ns[e] == o.n_steps, (t, "steps", int(ns[e]), o.n_steps)
        steps += o.n_steps
        if o.status > 3:
            continue  # an escaping exception: the reference has no state to compare
        assert mk[e] == o.time, (t, "time", int(mk[e]), o.time)
        assert ret[e] == o.ret, (t, "return")
        want = o.digest() if record else mask_done_ops(o.digest(), li)
        assert np.array_equal(batch.get_state(e), want), (t, "digest")
    batch.close()
    return steps


def check_instance_steps(eng, torch, li, cfg, sub, B=6, L=600, tag="", tape=None):
    """Step mode of the THROUGHPUT build against the oracle: every env.step of every env."""
    tape = make_tape(sub, B, L) if tape is None else tape
    batch = eng.Batch(li, B, **cfg)
    oras = [op.OracleEnv(li, **cfg) for _ in range(B)]
    out = batch.reset()
    st = out.status.cpu().numpy()
    t0 = f"{tag} (J={li.n_jobs}, M={li.n_machines}, T={li.n_transports}, step mode)"
    for e, o in enumerate(oras):
        assert int(st[e]) == o.status, (t0, "reset status", e, int(st[e]), o.status)
    if oras[0].status > 3:  # the reference's constructor raised
        batch.close()
        return 0
    obs = out.obs.cpu().numpy()
    for e, o in enumerate(oras):
        assert obs[e].tobytes() == o.obs(), (t0, "reset obs", e)
    dev_tape = torch.from_numpy(np.ascontiguousarray(tape.T)).cuda()
    steps = 0
    live = [True] * B
    for i in range(L):
        if not any(live):
            break
        out = batch.step(dev_tape[i])
        st = out.status.cpu().numpy(); rew = out.reward.cpu().numpy(); tm = out.time.cpu().numpy()
        term = out.terminated.cpu().numpy(); trunc = out.truncated.cpu().numpy()
        cand = out.cand.cpu().numpy(); obs = out.obs.cpu().numpy(); mk = out.makespan.cpu().numpy()
        for e, o in enumerate(oras):
            if not live[e]:
                continue
            o.step(int(tape[e, i]))
            t = f"{t0} env {e} step {i}"
            assert int(st[e]) == o.status, (t, "status", int(st[e]), o.status)
            steps += 1
            if o.status > 3:
                live[e] = False  # an escaping exception: nothing more to compare, the env is dead in both
                continue
            assert rew[e] == o.reward, (t, "reward", rew[e], o.reward)
            assert int(tm[e]) == o.time, (t, "time")
            assert bool(term[e]) == o.terminated and bool(trunc[e]) == o.truncated, (t, "flags")
            assert int(mk[e]) == (o.time if o.terminated else -1), (t, "makespan")
            oc = o.cand()
            if o.status == 3:  # success=False: possible_transitions is empty (state.py:202-210)
                assert int(cand[e, 3]) == 0, (t, "candidates after a failed step")
            else:
                assert [int(x) for x in cand[e, :3]] == [int(x) for x in oc[:3]] and \
                    int(cand[e, 3]) == min(int(oc[3]), 32767), (t, "candidate", cand[e].tolist(), oc.tolist())
            assert obs[e].tobytes() == o.obs(), (t, "obs")
            if o.status != 0:
                live[e] = False
    batch.close()
    return steps


def main(n=200, seed=0, jrange=(1, 8), step_every=1):
    import torch
    from jobshoplab_b200 import engine as eng

    rng = np.random.default_rng(seed)
    total = total_fast = total_step = 0
    for i in range(n):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        li, cfg = fuzz_engine.random_instance(sub, simple=(None, 3, 1, 2, 4)[i % 5], jrange=jrange)
        tape = make_tape(sub, 6, 600)
        tag = f"gpu fuzz instance {i} seed {seed}"
        total += check_instance(eng, torch, li, cfg, sub, tag=tag, tape=tape)
        total_fast += check_instance(eng, torch, li, cfg, sub, tag=tag, tape=tape, record=False)
        if i % step_every == 0:
            total_step += check_instance_steps(eng, torch, li, cfg, sub, tag=tag, tape=tape)
    print(f"gpu fuzz ok: {n} instances; env-steps compared: {total} (logging build, rollout), {total_fast} "
          f"(throughput build, rollout), {total_step} (throughput build, step mode, every step)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
