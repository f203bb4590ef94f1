"""BASELINE.json's full size on one GPU (configs[2]: ta41-shape batch, B = 1 000 000 instances) through
size-independent properties: every episode terminates, lower bound <= makespan <= max_allowed_time, step and
return bookkeeping, run-to-run determinism, shard invariance (a job split over two batches reproduces the single
batch), and a sample of envs against the CPU oracle on the tables read back from the device.  The instances are
drawn on the device (jsl_batch_randomize), so the test needs no host-side generator."""
import dataclasses
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _batch(li, B, seed, env_base=0):
    from jobshoplab_b200 import engine

    tile = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.int32), (B, len(a))))
    b = engine.Batch(li, B, seed=seed, variants=(tile(li.op_machine), tile(li.op_duration),
                                                 np.full(B, li.lower_bound, np.int32),
                                                 np.full(B, li.max_allowed_time, np.int32)))
    b.set_env_base(env_base)
    b.randomize(seed=seed + 1, min_duration=1, max_duration=100)
    return b


def test_million_env_rollout_properties():
    import bench
    from jobshoplab_b200 import abi
    from oracle import oracle_py as op

    li = bench.template_instance()
    B, N, J = 1_000_000, li.n_ops, li.n_jobs
    b = _batch(li, B, seed=11)
    r = b.rollout("random", reset_first=True, writeback=False)
    mk, ns = r.makespan.clone(), r.n_steps.clone()
    ret, st, noop = r.ret.clone(), r.status.clone(), r.no_op_count.clone()
    assert bool((st == abi.ST_DONE).all())
    # one accepted transition per operation, every other step is a no-op
    assert bool((ns - noop == N).all()) and int(ns.min()) >= N
    # bounds of every instance (read back in slices), rewards: terminal reward in (0, 1], penalties <= 0
    for first in (0, B // 2, B - 50_000):
        _, _, od, lb, mat = b.get_variants(first, 50_000)
        m = mk[first:first + 50_000].cpu().numpy()
        assert (lb <= m).all() and (m <= mat).all() and (od >= 1).all() and (od <= 99).all()
        assert (mat == od.sum(1)).all()
    assert float(ret.max()) <= 1.0 and float(ret.min()) > -1.0
    # determinism
    r2 = b.rollout("random", reset_first=True, writeback=False)
    assert torch.equal(r2.makespan, mk) and torch.equal(r2.n_steps, ns) and torch.equal(r2.ret, ret)
    # a sample of envs against the CPU oracle on the tables the device generated
    rng = np.random.default_rng(0)
    for e in rng.integers(0, B, 24).tolist():
        om, ot, od, lb, mat = b.get_variants(e, 1)
        liv = dataclasses.replace(li, op_machine=om[0], op_tool=ot[0], op_duration=od[0], lower_bound=int(lb[0]),
                                  max_allowed_time=int(mat[0]))
        ref = op.rollout_batch([liv], 1, abi.POLICY_RANDOM, seed=11, env_base=e)
        assert int(ref["makespan"][0]) == int(mk[e]) and int(ref["n_steps"][0]) == int(ns[e]), e
        assert float(ref["ret"][0]) == float(ret[e]), e
    b.close()
    # shard invariance: two half batches with their global offsets == the corresponding slices
    H = 100_000
    full = _batch(li, 2 * H, seed=11)
    rf = full.rollout("random", reset_first=True, writeback=False)
    fm, fn = rf.makespan.clone(), rf.n_steps.clone()
    full.close()
    for k in range(2):
        half = _batch(li, H, seed=11, env_base=k * H)
        rh = half.rollout("random", reset_first=True, writeback=False)
        assert torch.equal(rh.makespan, fm[k * H:(k + 1) * H]) and torch.equal(rh.n_steps, fn[k * H:(k + 1) * H])
        half.close()
    assert torch.equal(fm, mk[:2 * H]) and torch.equal(fn, ns[:2 * H])  # a prefix of the million-env job


def test_step_mode_equals_rollout_at_size():
    """65 536 different instances stepped one env.step at a time with random device actions; the recorded action
    tape replayed by the fused rollout kernel must end in the same makespans / step counts / returns."""
    import bench
    from jobshoplab_b200 import abi

    li = bench.template_instance()
    B, T = 65_536, 1700
    b = _batch(li, B, seed=5)
    b.reset()
    tape = torch.full((B, T), 2, dtype=torch.uint8, device="cuda")  # 2 = out of the action space: stops a replay
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    mk = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    ns = torch.zeros(B, dtype=torch.int32, device="cuda")
    ret = torch.zeros(B, dtype=torch.float64, device="cuda")
    alive = torch.ones(B, dtype=torch.bool, device="cuda")
    acts = torch.empty(B, dtype=torch.uint8, device="cuda")
    for n in range(T):
        acts.random_(0, 2, generator=g)
        tape[alive, n] = acts[alive]
        out = b.step(acts)
        ns += alive.to(torch.int32)
        ret += torch.where(alive, out.reward, torch.zeros_like(out.reward))
        fin = alive & (out.terminated != 0)
        mk[fin] = out.makespan[fin]
        alive &= ~fin
        if not bool(alive.any()):
            break
    assert not bool(alive.any())
    _, _, _, lb, mat = b.get_variants(0, B)
    d = mk.cpu().numpy()
    assert (lb <= d).all() and (d <= mat).all()
    b2 = _batch(li, B, seed=5)  # the same instances
    b2.reset()
    r = b2.rollout("tape", tape=tape)
    assert torch.equal(r.makespan, mk) and torch.equal(r.n_steps, ns) and torch.equal(r.ret, ret)
    assert bool((r.status == abi.ST_DONE).all())
    b.close(); b2.close()
