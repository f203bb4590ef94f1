"""GPU parity: the CUDA engine, driven through the C ABI (jobshoplab_b200.engine -> libjsl_b200.so),
against the golden traces of the unmodified reference and against the CPU oracle.

Bit-exact bar: canonical state digest (incl. candidate list), reward (f64), flags, packed observation.
Run on the B200 box with `pytest -m gpu`.
"""
import numpy as np
import pytest

import _golden as G
from jobshoplab_b200 import abi

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _engine():
    from jobshoplab_b200 import engine
    return engine


EXC_TO_STATUS = {"BufferFullError": abi.ST_BUFFER_FULL, "UnsuccessfulStateMachineResult": abi.ST_UNSUCCESSFUL,
                 "InvalidValue": abi.ST_NO_POSSIBLE, "NotImplementedError": abi.ST_OTHER_INVALID}


def step_replay(sc: G.Scenario, B=5, record=True):
    """Step mode: every env of a small batch gets the golden tape; env 0 and env B-1 are checked each step.
    record=True runs the logging build of the library (the digest of DONE operations needs the op log);
    record=False runs the throughput build -- the one bench.py measures -- and checks everything but the digest."""
    eng = _engine()
    batch = eng.Batch(sc.lowered, B, record_ops=record, **sc.cfg)
    out = batch.reset()
    torch.cuda.synchronize()
    probe = (0, B - 1)
    for e in probe:
        assert not record or G.hash_i32(batch.get_state(e)) == int(sc.state_hash[0]), "reset state"
    obs = out.obs.cpu().numpy()
    for e in probe:
        assert G.hash_bytes(obs[e].tobytes()) == int(sc.obs_hash[0]), "reset obs"
    n = sc.meta["n_steps"]
    acts = torch.zeros(B, dtype=torch.uint8, device="cuda")
    for i in range(n):
        acts.fill_(int(sc.actions[i]))
        out = batch.step(acts)
        st = out.status.cpu().numpy()
        assert (st <= abi.ST_STEP_FAILED).all(), (i, st)
        rew = out.reward.cpu().numpy(); tm = out.time.cpu().numpy()
        term = out.terminated.cpu().numpy(); trunc = out.truncated.cpu().numpy()
        obs = out.obs.cpu().numpy(); cand = out.cand.cpu().numpy()
        for e in probe:
            assert not record or G.hash_i32(batch.get_state(e)) == int(sc.state_hash[i + 1]), f"state after step {i}"
            assert rew[e] == float(sc.reward[i]), f"reward at step {i}"
            assert tm[e] == int(sc.time[i + 1])
            assert (int(term[e]) | (int(trunc[e]) << 1)) == int(sc.flags[i])
            assert int(cand[e, 3]) == min(int(sc.n_possible[i + 1]), 32767)
            assert G.hash_bytes(obs[e].tobytes()) == int(sc.obs_hash[i + 1]), f"obs after step {i}"
    if sc.exception:
        acts.fill_(int(sc.actions[n]))
        st = batch.step(acts).status.cpu().numpy()
        assert (st == EXC_TO_STATUS[sc.exception]).all(), (sc.exception, st)
    else:
        assert not record or np.array_equal(batch.get_state(0), sc.final_digest)
        mk = batch.out.makespan.cpu().numpy()
        terminated = bool(int(sc.flags[-1]) & 1) if n else False
        assert (mk == (sc.meta["makespan"] if terminated else -1)).all()  # info["makespan"], env.py:408
        if not sc.cut:
            acts.fill_(1)
            st = batch.step(acts).status.cpu().numpy()
            assert (st == abi.ST_ENV_DONE).all()  # env.py:441
    batch.close()


def _deterministic(group):
    return [n for n in G.names(group) if G.Scenario(group, n).lowered.is_deterministic]


SHORT = {"classic": ["2x2/fifo", "3x3/random0", "ft06/fifo", "ft06/spt", "ft06/random0", "ft10/mwkr", "la16/random1",
                     "ft20/spt", "ta01/random0", "ta41/fifo", "swv16/fifo", "yn1/random0"],
         "solutions": ["3x3/solution", "ft06/solution", "ft10/solution", "la01/solution", "ta01/solution"]}


@pytest.mark.parametrize("group", ["classic", "solutions", "transport", "buffers", "features", "truncation", "oparray", "tassel", "big"])
def test_step_mode_matches_reference_traces(group):
    if group in SHORT:
        names = SHORT[group]
    else:
        names = _deterministic(group)
        if group == "transport":
            names = [n for n in names if not n.startswith(("ta41", "la21"))] + ["ta41-t/random0", "la21-t/noearly/spt"]
        if group == "buffers":
            names = names[::3]
    for name in names:
        try:
            step_replay(G.Scenario(group, name))
        except AssertionError as ex:
            raise AssertionError(f"{group}:{name}: {ex}") from ex


@pytest.mark.parametrize("group", ["classic", "solutions", "transport", "buffers", "features", "truncation", "oparray", "tassel"])
def test_throughput_build_step_mode(group):
    """The same traces through libjsl_b200.so proper (no logs compiled in): reward, time, flags, candidate count
    and the observation of every step."""
    names = SHORT.get(group) or _deterministic(group)
    if group in ("transport", "buffers", "features"):
        names = names[::5]
    for name in names:
        try:
            step_replay(G.Scenario(group, name), record=False)
        except AssertionError as ex:
            raise AssertionError(f"{group}:{name}: {ex}") from ex


@pytest.mark.parametrize("level", ["0", "1"])
def test_specialisation_levels_agree_on_gpu(level, monkeypatch):
    """libjsl_b200.so picks the most specialised kernel an instance allows (classic: level 2, transport: level 1,
    ordered / bounded buffers: level 3).  JSL_MAX_LEVEL (read by jsl_batch_create) forces the less specialised
    kernels, which must reproduce the same reference traces -- otherwise the general code would only ever see the
    setup-time / outage instances."""
    monkeypatch.setenv("JSL_MAX_LEVEL", level)
    picks = [("classic", n) for n in SHORT["classic"][:5]] + [("transport", n) for n in _deterministic("transport")[::6]]
    if level == "0":
        picks += [("buffers", n) for n in _deterministic("buffers")[1::3]] + [("features", n) for n in _deterministic("features")[::3]]
    for group, name in picks:
        try:
            step_replay(G.Scenario(group, name), record=False)
        except AssertionError as ex:
            raise AssertionError(f"level {level} {group}:{name}: {ex}") from ex


@pytest.mark.parametrize("group", ["classic", "solutions", "transport", "buffers", "features", "truncation", "big"])
def test_rollout_tape_matches_reference(group):
    """Rollout mode (state resident in shared memory for the whole episode): every deterministic golden
    scenario of an instance becomes one env of a TAPE-policy batch; final makespan / steps / return /
    status / no-op count / final digest must equal the reference's."""
    eng = _engine()
    by_inst = {}
    for name in _deterministic(group):
        sc = G.Scenario(group, name)
        key = (sc.source_text, tuple(sorted(sc.cfg.items())))
        by_inst.setdefault(key, []).append(sc)
    for scs in by_inst.values():
        B = len(scs)
        L = max(len(s.actions) for s in scs) + 1
        tape = np.full((B, L), 2, dtype=np.uint8)  # past its tape an env gets an out-of-space action and stops
        for i, s in enumerate(scs):
            tape[i, : len(s.actions)] = s.actions
        for record in (True, False):  # logging build (+ final digest) and throughput build
            batch = eng.Batch(scs[0].lowered, B, record_ops=record, **scs[0].cfg)
            batch.reset()
            r = batch.rollout("tape", tape=torch.from_numpy(tape).cuda())
            mk, ns = r.makespan.cpu().numpy(), r.n_steps.cpu().numpy()
            ret, st, noop = r.ret.cpu().numpy(), r.status.cpu().numpy(), r.no_op_count.cpu().numpy()
            for i, s in enumerate(scs):
                tag = f"{group}:{s.name}"
                if s.exception:
                    assert st[i] == EXC_TO_STATUS[s.exception], tag
                    continue
                assert ns[i] == s.meta["n_steps"], tag
                assert mk[i] == s.meta["makespan"], tag
                assert noop[i] == s.meta["no_op_count"], tag
                assert ret[i] == float(np.sum(s.reward[: 0]) + sum(float(x) for x in s.reward)), tag
                assert st[i] in (abi.ST_DONE, abi.ST_TRUNCATED) or (s.cut and st[i] == abi.ST_ACTION_OUT_OF_SPACE), tag
                assert not record or np.array_equal(batch.get_state(i), s.final_digest), tag
            batch.close()


def test_rollout_policies_match_reference_rules():
    """In-kernel FIFO / SPT / MWKR / RANDOM policies (dipatching_benchmark.py:17-67) reproduce the
    makespans and step counts the reference's own rule drivers produced."""
    eng = _engine()
    for group in ("classic", "transport"):
        for name in G.names(group):
            pol = name.rsplit("/", 1)[1]
            sc = G.Scenario(group, name)
            B = 7
            batch = eng.Batch(sc.lowered, B, seed=sc.meta.get("seed", 0), **sc.cfg)
            if pol.startswith("random"):
                batch.set_env_base(sc.meta.get("env_idx", 0) - 3)  # env 3 draws the golden Philox stream
                r = batch.rollout("random", reset_first=True)
                probe = [3]
            else:
                r = batch.rollout(pol, reset_first=True)
                probe = range(B)
            mk, ns = r.makespan.cpu().numpy(), r.n_steps.cpu().numpy()
            for e in probe:
                assert ns[e] == sc.meta["n_steps"] and mk[e] == sc.meta["makespan"], f"{group}:{name} env {e}"
            batch.close()


def test_wide_instances_with_queues_match_oracle():
    """Random instances with 33..100 jobs (2 and 4 lane words), FIFO / LIFO / DUMMY buffers with finite capacities,
    with and without setup times / outages (levels 3, 0 and 4), random action tapes: rollout mode against the CPU
    oracle -- final digest, makespan, steps, return, status.  The golden groups of that width are classic only."""
    import fuzz_engine
    import gpu_fuzz

    eng = _engine()
    rng = np.random.default_rng(77)
    for i in range(18):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        li, cfg = fuzz_engine.random_instance(sub, simple=(3, None, 4)[i % 3], jrange=((33, 64), (65, 101))[(i // 3) % 2])
        gpu_fuzz.check_instance(eng, torch, li, cfg, sub, B=6, L=1500, tag=f"wide instance {i}")


def test_gpu_differential_fuzz():
    """Small random instances of every specialisation level on the GPU against the oracle (tests/gpu_fuzz.py)."""
    import gpu_fuzz

    gpu_fuzz.main(100, 3)


def test_packed_record_equals_separate_arrays():
    """jsl_step_out.record (one 32-byte jsl_step_record per env, what engine.Batch uses) carries exactly what the
    separate per-field arrays of the ABI carry."""
    import ctypes as C

    eng = _engine()
    sc = G.Scenario("transport", "ft10-t/random0")
    B = 9
    a, b = eng.Batch(sc.lowered, B, **sc.cfg), eng.Batch(sc.lowered, B, **sc.cfg)
    dev = "cuda"
    leg = dict(reward=torch.zeros(B, dtype=torch.float64, device=dev), terminated=torch.zeros(B, dtype=torch.uint8, device=dev),
               truncated=torch.zeros(B, dtype=torch.uint8, device=dev), status=torch.zeros(B, dtype=torch.uint8, device=dev),
               time=torch.zeros(B, dtype=torch.int32, device=dev), makespan=torch.zeros(B, dtype=torch.int32, device=dev),
               cand=torch.zeros((B, 4), dtype=torch.int16, device=dev))
    obs2 = torch.zeros_like(b.out.obs)
    legacy = abi.StepOut(obs2.data_ptr(), leg["reward"].data_ptr(), leg["terminated"].data_ptr(), leg["truncated"].data_ptr(),
                         leg["status"].data_ptr(), leg["time"].data_ptr(), leg["makespan"].data_ptr(), leg["cand"].data_ptr(), None)
    a.reset()
    assert b.L.jsl_reset(b.h, None, legacy, b._stream()) == 0
    rng = np.random.default_rng(5)
    for i in range(len(sc.actions) + 2):
        acts = torch.from_numpy(rng.integers(0, 2, B).astype(np.uint8)).cuda()
        out = a.step(acts)
        assert b.L.jsl_step(b.h, acts.data_ptr(), legacy, b._stream()) == 0
        torch.cuda.synchronize()
        for k, v in leg.items():
            assert torch.equal(getattr(out, k).contiguous(), v), (i, k)
        assert torch.equal(out.obs, obs2), (i, "obs")
    a.close(); b.close()


def test_uploaded_tables_are_validated_on_the_device():
    """jsl_batch_stage/commit_variants used to pack whatever machine ids they were handed; a machine outside
    [0, M) is now neutralised while packing and reported by jsl_batch_check; wrong shapes never reach the library."""
    import bench
    from jobshoplab_b200.exceptions import InvalidValue

    eng = _engine()
    tmpl = bench.template_instance()
    om, od = bench.make_instances(8)
    lb, mat = bench.taillard_bounds(om, od)
    batch = eng.Batch(tmpl, 8, variants=(om, od, lb, mat))
    batch.check()
    with pytest.raises(InvalidValue):
        batch.load_variants(om[:4], od, lb, mat)
    with pytest.raises(InvalidValue):
        eng.Batch(tmpl, 8, variants=(om[:, :-1], od[:, :-1], lb, mat))
    bad = om.copy(); bad[3, 17] = 77
    batch.load_variants(bad, od, lb, mat)
    r = batch.rollout("random", reset_first=True)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == abi.ST_DONE).all()  # ran on the sanitised table instead of indexing out of bounds
    with pytest.raises(InvalidValue):
        batch.check()
    batch.close()



def test_maximum_sizes_and_limits():
    """The engine's limits: 128 jobs x 32 machines (4 lane words, every machine bit of the masks in
    use), 128 AGVs, and a ragged variant with 1..32 operations per job run bit-exactly against the oracle under three policies;
    one machine more is refused with an error, not a crash."""
    from jobshoplab_b200 import compiler as C
    from jobshoplab_b200.exceptions import InvalidValue
    from oracle import oracle_py as op

    eng = _engine()
    rng = np.random.default_rng(4)
    J, M = 128, 32
    lines = [f"{J} {M}"]
    for j in range(J):
        lines.append(" ".join(f"{m} {int(rng.integers(1, 50))}" for m in rng.permutation(M)))
    li = C.compile_spec_text("\n".join(lines) + "\n")
    assert (li.n_jobs, li.n_machines, li.n_transports) == (128, 32, 128)
    # ragged variant: job j keeps its first 1 + (7 j mod 32) operations
    import dataclasses
    keep = [1 + (7 * j) % 32 for j in range(J)]
    jos = np.concatenate([[0], np.cumsum(keep)]).astype(np.int32)
    sel = np.concatenate([np.arange(j * M, j * M + keep[j]) for j in range(J)])
    ragged = dataclasses.replace(li, job_op_start=jos, op_machine=li.op_machine[sel], op_tool=li.op_tool[sel],
                                 op_duration=li.op_duration[sel], lower_bound=1, max_allowed_time=int(li.op_duration[sel].sum()))
    for inst in (li, ragged):
        batch = eng.Batch(inst, 3, seed=9)
        for pol, code in (("fifo", abi.POLICY_FIFO), ("spt", abi.POLICY_SPT), ("random", abi.POLICY_RANDOM)):
            r = batch.rollout(pol, reset_first=True)
            o = op.OracleEnv(inst)
            o.rollout(code, seed=9, env_idx=1)
            assert int(r.status[1]) == abi.ST_DONE and o.terminated
            assert (int(r.makespan[1]), int(r.n_steps[1]), float(r.ret[1])) == (o.time, o.n_steps, o.ret), pol
        batch.close()
    lines = [f"4 33"] + [" ".join(f"{m} 3" for m in range(33)) for _ in range(4)]
    with pytest.raises(InvalidValue) as e:
        eng.Batch(C.compile_spec_text("\n".join(lines) + "\n"), 2)
    assert "32 machines" in str(e.value)


def test_packed_variant_upload_equals_int32_upload():
    """jsl_batch_stage_variants_packed (uint8 machines / uint16 durations, what bench.py's e2e path uploads) leaves the
    same tables on the device as the int32 path; values that do not fit are refused on the host."""
    import bench
    from jobshoplab_b200.exceptions import InvalidValue

    eng = _engine()
    tmpl = bench.template_instance()
    om, od = bench.make_instances(64, first=128)
    lb, mat = bench.taillard_bounds(om, od)
    a = eng.Batch(tmpl, 64, variants=(om, od, lb, mat))
    zeros = np.zeros_like(om)
    b = eng.Batch(tmpl, 64, variants=(zeros, zeros + 1, lb * 0 + 1, mat * 0 + 600))
    pm, pd = eng.Batch.pack_variants(om, od)
    b.stage_variants_packed(pm, pd, torch.from_numpy(lb).pin_memory(), torch.from_numpy(mat).pin_memory())
    b.commit_variants()
    b.check()
    for x, y in zip(a.get_variants(), b.get_variants()):
        assert np.array_equal(x, y)
    ra, rb = a.rollout("spt", reset_first=True), b.rollout("spt", reset_first=True)
    assert torch.equal(ra.makespan, rb.makespan) and torch.equal(ra.ret, rb.ret)
    with pytest.raises(InvalidValue):
        eng.Batch.pack_variants(om, od * 1000)
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("group,name", [("classic", "ft06/random0"), ("buffers", "la01-t/fifo1/random0"),
                                        ("features", "full_feature_3x3_instance/random0"), ("truncation", "ft06/trunc0/random0"),
                                        ("transport", "ft10-t/random0")])
def test_fused_auto_reset_equals_step_then_reset(group, name):
    """JSL_STEP_AUTO_RESET (the reset of a finished env inside the step launch) against the same thing done with two
    launches and a host decision: step, look at the flags, jsl_reset with a mask, keep reward / flags / makespan of
    the finished step.  Records, `ended`, observations and final states, over several episodes per env -- including
    envs that die with an escaped exception (capacity-1 buffers) and truncated ones."""
    eng = _engine()
    sc = G.Scenario(group, name)
    B = 48
    a, b = eng.Batch(sc.lowered, B, seed=3, **sc.cfg), eng.Batch(sc.lowered, B, seed=3, **sc.cfg)
    if not sc.lowered.is_deterministic:
        seeds = torch.from_numpy(np.tile(sc.start_seeds, (B, 1)).astype(np.int32)).cuda()
        a.set_start_seeds(seeds); b.set_start_seeds(seeds)
    a.reset(); b.reset()
    rng = np.random.default_rng(11)
    n_ended = 0
    kinds = set()
    for i in range(400):
        acts = torch.from_numpy((rng.random(B) < 0.7).astype(np.uint8)).cuda()
        fused = a.step(acts, auto_reset=True)
        out = b.step(acts)
        done = (out.terminated | out.truncated | (out.status >= abi.ST_STEP_FAILED).to(torch.uint8)).contiguous()
        want_ended = torch.where(done.bool(), out.status, torch.zeros_like(out.status))
        if bool(done.any()):
            final = b.record.clone()
            b.reset(done)
            rec = b.record
            rec[:, 0:8] = final[:, 0:8]; rec[:, 12:16] = final[:, 12:16]; rec[:, 24:26] = final[:, 24:26]
        n_ended += int(done.sum())
        kinds |= set(want_ended[want_ended > 0].cpu().tolist())
        assert torch.equal(fused.ended, want_ended), (i, fused.ended.cpu().tolist(), want_ended.cpu().tolist())
        assert torch.equal(a.record[:, :27], b.record[:, :27]), (i, "record")
        assert torch.equal(a.out.obs, b.out.obs), (i, "obs")
    assert n_ended >= B, n_ended  # every env finished at least one episode on average
    for e in (0, B // 2, B - 1):
        assert np.array_equal(a.get_state(e), b.get_state(e)), e
    # without the flag the byte stays 0, and unknown flag bits are refused
    assert int(b.out.ended.max()) == 0
    bad = abi.StepOut(a._step_out.obs, None, None, None, None, None, None, None, a.record.data_ptr(), 2, 0)
    assert a.L.jsl_step(a.h, acts.data_ptr(), bad, a._stream()) != 0
    a.close(); b.close()
    print(group, name, "episodes ended:", n_ended, "with statuses", sorted(kinds))
