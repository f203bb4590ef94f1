"""CPU-side check of the CUDA engine SOURCE: jobshoplab_b200/csrc/jsl_engine.cuh compiled with loop
primitives (tests/hostsim) must reproduce the oracle / the reference traces step by step.  This catches
logic errors in the kernel code without a GPU; the GPU parity tests (test_cuda_parity.py) then cover the
real warp execution.  The product never loads this build."""
import numpy as np
import pytest

import _golden as G
from hostsim.hostsim_py import HostSimEnv
from jobshoplab_b200 import abi

PICK = {
    "classic": ["2x2/fifo", "3x3/random0", "ft06/spt", "ft06/random1", "ft10/mwkr", "la16/random0", "ft20/fifo",
                "ta01/random0", "ta41/spt", "swv16/fifo", "yn1/random0"],
    "solutions": ["3x3/solution", "ft06/solution", "ft10/solution", "la16/solution", "ta01/solution", "dmu35/solution"],
    "transport": ["ft06-t/fifo", "ft06-t/noearly/random0", "ft10-t/spt", "la01-t/noearly/mwkr", "ft20-t/random1",
                  "ta41-t/noearly/random0"],
}
EXC_TO_STATUS = {"BufferFullError": abi.ST_BUFFER_FULL, "UnsuccessfulStateMachineResult": abi.ST_UNSUCCESSFUL,
                 "InvalidValue": abi.ST_NO_POSSIBLE, "NotImplementedError": abi.ST_OTHER_INVALID}


def replay(sc):
    e = HostSimEnv(sc.lowered, **sc.cfg)
    assert G.hash_i32(e.digest()) == int(sc.state_hash[0]), "reset"
    assert G.hash_bytes(e.obs()) == int(sc.obs_hash[0]), "reset obs"
    n = sc.meta["n_steps"]
    for i in range(n):
        st = e.step(int(sc.actions[i]))
        assert st <= abi.ST_STEP_FAILED
        assert G.hash_i32(e.digest()) == int(sc.state_hash[i + 1]), f"state after step {i}"
        assert e.reward == float(sc.reward[i]), f"reward at step {i}"
        assert (int(e.terminated) | (int(e.truncated) << 1)) == int(sc.flags[i])
        assert e.n_possible == int(sc.n_possible[i + 1])
        assert G.hash_bytes(e.obs()) == int(sc.obs_hash[i + 1]), f"obs after step {i}"
    if sc.exception:
        assert e.step(int(sc.actions[n])) == EXC_TO_STATUS[sc.exception]
    else:
        assert np.array_equal(e.digest(), sc.final_digest)
        assert e.noop_count == sc.meta["no_op_count"]


@pytest.mark.parametrize("group", ["classic", "solutions", "transport", "buffers", "features", "truncation", "oparray", "tassel", "big"])
def test_engine_source_matches_reference_traces(group):
    names = PICK.get(group)
    if names is None:
        names = [n for n in G.names(group) if G.Scenario(group, n).lowered.is_deterministic]
        if group == "buffers":
            names = names[::2]
    for name in names:
        try:
            replay(G.Scenario(group, name))
        except AssertionError as ex:
            raise AssertionError(f"{group}:{name}: {ex}") from ex


@pytest.mark.parametrize("level", ["0", "1"])
def test_specialisation_levels_agree(level, monkeypatch):
    """Classic instances are normally served by the level-2 code (fast paths), transport instances by level 1,
    instances with ordered / bounded buffers but no setup times or outages by level 3: forced down to the simple /
    general code they must reproduce the same reference traces."""
    monkeypatch.setenv("JSL_HOSTSIM_MAX_LEVEL", level)
    for group in ("classic", "transport"):
        for name in PICK[group][:6]:
            try:
                replay(G.Scenario(group, name))
            except AssertionError as ex:
                raise AssertionError(f"level {level} {group}:{name}: {ex}") from ex
    if level == "0":  # buffers runs at level 3 by default (the other half of what the main test replays), features at 4
        for group, sl in (("buffers", slice(1, None, 2)), ("features", slice(0, None, 2))):
            names = [n for n in G.names(group) if G.Scenario(group, n).lowered.is_deterministic]
            for name in names[sl]:
                try:
                    replay(G.Scenario(group, name))
                except AssertionError as ex:
                    raise AssertionError(f"level 0 {group}:{name}: {ex}") from ex


def test_engine_policies():
    for name in ("ft06/spt", "ft10/mwkr", "la16/spt", "ta01/fifo", "ft06/random2"):
        sc = G.Scenario("classic", name)
        pol = name.rsplit("/", 1)[1]
        e = HostSimEnv(sc.lowered, seed=sc.meta.get("seed", 0))
        if pol.startswith("random"):
            e.rollout(abi.POLICY_RANDOM, env_idx=sc.meta["env_idx"])
        else:
            e.rollout(abi.POLICIES[pol])
        assert e.n_steps == sc.meta["n_steps"] and e.time == sc.meta["makespan"], name


def test_engine_source_differential_fuzz():
    """Random small instances (buffers, capacities, travel, setup, outages, truncation): engine source vs
    oracle, step by step -- tests/fuzz_engine.py."""
    import fuzz_engine

    from hostsim.hostsim_py import fused, invariants, set_fuse
    before = fused()
    fuzz_engine.main(500, 3)
    checks, violations = invariants()
    assert checks > 1000 and violations == 0
    # the chain fusions of the general event loop all fired in this run (they are what the kernels execute) ...
    fired = [b - a for a, b in zip(before, fused())]
    assert fired[0] > 100 and fired[2] > 1000 and fired[3] > 100 and fired[4] > 100, fired
    # ... and with every fusion off (one reference batch per loop iteration) the engine agrees with the oracle as well
    set_fuse(0)
    try:
        fuzz_engine.main(150, 4)
    finally:
        set_fuse(-1)


def test_engine_source_wide_instances():
    """33..100 jobs (2 and 4 lane words) with ordered / bounded buffers, setup times and outages: the golden
    groups only hold classic instances of that width."""
    import fuzz_engine

    rng = np.random.default_rng(77)
    for i in range(24):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        try:
            fuzz_engine.run_one(sub, max_steps=600, simple=(3, None, 4)[i % 3], jrange=((33, 64), (65, 101))[(i // 3) % 2])
        except AssertionError as ex:
            raise AssertionError(f"wide instance {i}: {str(ex)[:2000]}") from ex


def test_reset_of_a_used_env_equals_a_fresh_one():
    """What JSL_STEP_AUTO_RESET relies on: the engine's reset (env_init + the dummy step) does not depend on what the
    env's slot held before -- after a finished, a dead or an interrupted episode the state, observation and candidate
    equal those of a freshly created env.  Engine source on the CPU, every specialisation level incl. stochastic."""
    from hostsim.hostsim_py import HostSimEnv

    cases = [("classic", "ft06/random0"), ("transport", "ft10-t/random0"), ("buffers", "la01-t/fifo1/random0"),
             ("buffers", "la01-t/lifo2/noearly/random0"), ("features", "full_feature_3x3_instance/random0"),
             ("truncation", "ft06/trunc0/random0")]
    rng = np.random.default_rng(9)
    for group, name in cases:
        sc = G.Scenario(group, name)

        def make():
            e = HostSimEnv(sc.lowered, **sc.cfg)
            if not sc.lowered.is_deterministic:
                e.set_start_seeds(sc.start_seeds)
            return e

        fresh = make(); fresh.reset()
        want = (fresh.digest().tobytes(), fresh.obs())
        used = make(); used.reset()
        for stop_after in (7, 10**9, 23):  # interrupted, run to its end (or death), interrupted again
            n = 0
            while n < stop_after:
                st = used.step(int(rng.random() < 0.7))
                n += 1
                if st != 0:
                    break
            used.reset()
            assert (used.digest().tobytes(), used.obs()) == want, (group, name, stop_after)
