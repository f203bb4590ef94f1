"""The CPU oracle (oracle/jsl_oracle.c) pinned against traces of the UNMODIFIED reference.

Every golden scenario (tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference) is
replayed action by action; after each step the oracle's canonical state digest hash, packed
observation hash, reward (bit-exact f64), time, terminated/truncated flags and candidate count must
equal the reference's.  This includes the reference's own golden vectors: the 54 optimal-solution
replays of tests/end_to_end_tests/test_env_end_to_end.py:72-105 (group "solutions").
"""
import numpy as np
import pytest

import _golden as G
from jobshoplab_b200 import abi
from oracle import oracle_py as op

EXC_TO_STATUS = {
    "BufferFullError": abi.ST_BUFFER_FULL,
    "UnsuccessfulStateMachineResult": abi.ST_UNSUCCESSFUL,
    "InvalidValue": abi.ST_NO_POSSIBLE,
    "NotImplementedError": abi.ST_OTHER_INVALID,
}


def replay(sc: G.Scenario):
    env = op.OracleEnv.__new__(op.OracleEnv)
    op.OracleEnv.__init__(env, sc.lowered, **sc.cfg)
    if len(sc.samples):
        env.set_tape(sc.samples)
        env.reset()
    obs = {2: env.obs_oparray, 3: env.obs_tassel}.get(sc.cfg.get("obs_kind"), env.obs)
    assert G.hash_i32(env.digest()) == int(sc.state_hash[0]), "reset state"
    assert G.hash_bytes(obs()) == int(sc.obs_hash[0]), "reset obs"
    n = sc.meta["n_steps"]
    for i in range(n):
        st = env.step(int(sc.actions[i]))
        assert st <= abi.ST_STEP_FAILED, (i, abi.STATUS_NAMES[st])
        assert G.hash_i32(env.digest()) == int(sc.state_hash[i + 1]), f"state after step {i}"
        assert env.time == int(sc.time[i + 1])
        assert env.reward == float(sc.reward[i]), f"reward at step {i}"
        assert (int(env.terminated) | (int(env.truncated) << 1)) == int(sc.flags[i])
        assert env.n_possible == int(sc.n_possible[i + 1])
        assert G.hash_bytes(obs()) == int(sc.obs_hash[i + 1]), f"obs after step {i}"
    if sc.exception:
        st = env.step(int(sc.actions[n]))
        assert st == EXC_TO_STATUS[sc.exception], (sc.exception, abi.STATUS_NAMES[st])
    else:
        assert np.array_equal(env.digest(), sc.final_digest)
        assert env.time == sc.meta["makespan"]
        assert env.noop_count == sc.meta["no_op_count"]
        if not sc.cut:
            assert env.done
            assert env.step(1) == abi.ST_ENV_DONE  # env.py:441
    assert env.L.jslo_tape_pos(env.h) == len(sc.samples)


@pytest.mark.parametrize("group", G.GROUPS)
def test_oracle_matches_reference_traces(group):
    for name in G.names(group):
        try:
            replay(G.Scenario(group, name))
        except AssertionError as ex:
            raise AssertionError(f"{group}:{name}: {ex}") from ex


def test_oracle_policies_reproduce_rule_rollouts():
    """fifo/spt/mwkr chosen by the oracle's own policy code == the tape the reference's rules produced."""
    for group in ("classic", "transport"):
        for name in G.names(group):
            pol = name.rsplit("/", 1)[1]
            if pol not in ("fifo", "spt", "mwkr"):
                continue
            sc = G.Scenario(group, name)
            env = op.OracleEnv(sc.lowered, **sc.cfg)
            n = env.rollout(abi.POLICIES[pol])
            assert n == sc.meta["n_steps"], name
            assert env.time == sc.meta["makespan"], name
            assert np.array_equal(env.digest(), sc.final_digest), name


def test_oracle_random_policy_is_the_philox_tape():
    for name in G.names("classic"):
        if "random" not in name:
            continue
        sc = G.Scenario("classic", name)
        env = op.OracleEnv(sc.lowered, **sc.cfg)
        env.rollout(abi.POLICY_RANDOM, seed=sc.meta["seed"], env_idx=sc.meta.get("env_idx", 0))
        assert env.n_steps == sc.meta["n_steps"] and env.time == sc.meta["makespan"], name


def test_published_makespans():
    """rule-rollout makespans and a few optimal-solution replays."""
    expect = {"ft06/fifo": 68, "ft06/spt": 88, "ft06/mwkr": 83, "ft10/fifo": 1262, "ft10/spt": 1074,
              "ft10/mwkr": 1262, "la16/fifo": 1230, "la16/spt": 1156, "la16/mwkr": 1268, "ta01/fifo": 1830,
              "ta01/spt": 1462, "ta01/mwkr": 1501, "ta41/fifo": 2973, "ta41/spt": 2465, "ta41/mwkr": 3120}
    for name, mk in expect.items():
        assert G.Scenario("classic", name).meta["makespan"] == mk, name
    expect_t = {"ft06-t/fifo": 83, "ft06-t/spt": 85, "ft06-t/mwkr": 85, "ft10-t/fifo": 1463, "ft10-t/spt": 1167,
                "ft10-t/mwkr": 1286, "ft20-t/fifo": 1591, "la01-t/spt": 883, "la02-t/mwkr": 953}
    for name, mk in expect_t.items():
        assert G.Scenario("transport", name).meta["makespan"] == mk, name
    sol = {"3x3": 11, "ft06": 55, "ft10": 930, "ft20": 1165, "la01": 666, "ta01": 1231, "dmu32": 5927, "swv20": 2823}
    for name, mk in sol.items():
        assert G.Scenario("solutions", f"{name}/solution").meta["makespan"] == mk, name
