"""JobShopLabEnv / BatchedJobShopLabEnv on the GPU, used the way the reference's own tests use
jobshoplab.JobShopLabEnv (tests/end_to_end_tests/test_env_end_to_end.py, tests/unit_tests/test_env.py)."""
import random

import numpy as np
import pytest

import _golden as G

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _env(group, name, **cfg_over):
    from jobshoplab_b200 import Config, JobShopLabEnv
    from jobshoplab_b200 import compiler as C

    sc = G.Scenario(group, name)
    cfg = Config()
    for k, v in cfg_over.items():
        if k == "allow_early_transport":
            cfg.state_machine.allow_early_transport = v
    repo = C.DslStrRepository(sc.source_text) if sc.kind != "spec" else None
    comp = C.Compiler(cfg, repo=repo) if repo else C.compile_spec_text(sc.source_text)
    return JobShopLabEnv(config=cfg, compiler=comp), sc


def test_optimal_solution_replay_protocol():
    """test_env_end_to_end.py:72-105 against the env API, with the solution sequence recovered from the
    golden tape's accepted (machine, job, time) triples."""
    env, sc = _env("solutions", "ft06/solution")
    terminated = False
    for a in sc.actions.tolist():
        ev = env.state.possible_transitions[0]
        assert ev.component_id.startswith(("m-", "t-")) and ev.job_id.startswith("j-")
        obs, reward, terminated, truncated, info = env.step(int(a))
    assert terminated and not truncated
    assert env.state.state.time.time == 55 == info["makespan"]  # ft06 optimum, BASELINE.md 2b
    assert info["no_op_count"] == sc.meta["no_op_count"]
    from jobshoplab_b200.exceptions import EnvDone
    with pytest.raises(EnvDone):
        env.step(1)
    log = env.op_log()
    assert log.shape == (36, 2) and (log[:, 1] > log[:, 0]).all() and log[:, 1].max() == 55
    env.close()


def test_random_episodes_terminate_and_reset():
    """test_env_end_to_end.py:42-55: random actions, terminated != truncated, reset, repeat."""
    env, _ = _env("classic", "3x3/fifo")
    rng = random.Random(0)
    for _ in range(5):
        while not env.done:
            obs, r, term, trunc, info = env.step(rng.randint(0, 1))
        assert term != trunc
        assert set(obs) == {"job_running", "job_executed_on_machine", "job_progression", "machine_running",
                            "machine_progression", "available_jobs", "current_time", "current_transition"}
        assert obs["current_transition"].tolist() == [1.0, 1.0, 1.0]
        env.reset()
        assert not env.done
    env.close()


def test_buffer_full_escapes_step_like_the_reference():
    from jobshoplab_b200.exceptions import BufferFullError

    name = next(n for n in G.names("buffers") if G.Scenario("buffers", n).exception == "BufferFullError")
    env, sc = _env("buffers", name, allow_early_transport=sc_cfg(name))
    with pytest.raises(BufferFullError):
        for a in sc.actions.tolist():
            env.step(int(a))
    env.close()


def sc_cfg(name):
    return G.Scenario("buffers", name).cfg.get("allow_early_transport", True)


def test_batched_env_auto_reset_and_views():
    from jobshoplab_b200 import BatchedJobShopLabEnv, Config

    sc = G.Scenario("classic", "ft06/fifo")
    B = 257
    env = BatchedJobShopLabEnv(Config(), num_envs=B, instances=sc.lowered, auto_reset=True)
    obs, out = env.reset()
    assert obs["job_executed_on_machine"].shape == (B, 6, 6) and obs["current_time"].shape == (B, 1)
    acts = torch.ones(B, dtype=torch.uint8, device="cuda")
    finished = 0
    for _ in range(36 * 3):
        obs, rew, term, trunc, out = env.step(acts)
        finished += int(term.sum())
        mk = out.makespan[term.bool()]
        assert (mk == 68).all()  # ft06 always-accept makespan, BASELINE.md 2a
    assert finished == 3 * B
    assert int(out.status.max()) <= 1
    env.step_async(acts)
    obs2, rew2, term2, trunc2, _ = env.step_wait()
    assert obs2["job_progression"].shape == (B, 6) and env.single_action_space.n == 2
    r = env.rollout("mwkr", reset_first=True)
    assert (r.makespan == 83).all() and (r.n_steps == 46).all()
    assert env.state(5).state.time.time == 83
    env.close()


def test_multi_instance_batch_and_variants():
    """Several same-shaped instances in one batch: env b runs instance b % n_inst."""
    from jobshoplab_b200 import engine

    a = G.Scenario("classic", "ft10/spt")
    b = G.Scenario("classic", "la16/spt")
    batch = engine.Batch([a.lowered, b.lowered], 10)
    r = batch.rollout("spt", reset_first=True)
    mk = r.makespan.cpu().numpy()
    assert (mk[0::2] == 1074).all() and (mk[1::2] == 1156).all()
    om = np.stack([a.lowered.op_machine, b.lowered.op_machine, a.lowered.op_machine])
    od = np.stack([a.lowered.op_duration, b.lowered.op_duration, a.lowered.op_duration])
    lb = np.array([a.lowered.lower_bound, b.lowered.lower_bound, a.lowered.lower_bound], dtype=np.int32)
    mat = np.array([a.lowered.max_allowed_time, b.lowered.max_allowed_time, a.lowered.max_allowed_time], dtype=np.int32)
    v = engine.Batch(a.lowered, 9, variants=(om, od, lb, mat))
    mk = v.rollout("fifo", reset_first=True).makespan.cpu().numpy()
    assert mk.tolist() == [1262, 1230, 1262] * 3
    v.load_variants(torch.from_numpy(od * 0 + om).pin_memory(), torch.from_numpy(od[::-1].copy()).pin_memory(),
                    torch.from_numpy(lb), torch.from_numpy(mat))
    batch.close(); v.close()


def test_render_data_matches_reference_dashboard_model():
    """env.gantt_data() == DashboardDataMapper.map_states_to_gant_data(env.history) of the reference
    (tests/golden/gantt.json), through the public env; render() prints one line per bar."""
    import json
    from pathlib import Path

    gold = json.loads((Path(__file__).parent / "golden" / "gantt.json").read_text())
    for key in ("transport:ft06-t/random1", "classic:ft06/random0"):
        group, name = key.split(":")
        env, sc = _env(group, name)
        for a in sc.actions[: gold[key]["final"]["steps"]].tolist():
            env.step(int(a))
        data = env.gantt_data()
        want = gold[key]["final"]["data"]
        assert [dict(x) for x in data["schedules"]] == want["schedules"]
        assert [dict(x) for x in data["transports"]] == want["transports"]
        text = env.render()
        assert len(text.splitlines()) == len(want["schedules"]) + len(want["transports"])
        env.close()


def test_sb3_style_vec_env_matches_single_env():
    """the stable_baselines3 VecEnv protocol over B GPU envs; env 3 of the vector env is compared with
    the single drop-in env fed the same actions, across auto-resets."""
    from jobshoplab_b200 import Config
    from jobshoplab_b200 import compiler as C
    from jobshoplab_b200.vec_env import SB3VecEnv

    sc = G.Scenario("classic", "ft06/fifo")
    li = C.compile_spec_text(sc.source_text)
    B = 16
    venv = SB3VecEnv(Config(), num_envs=B, compiler=li, seed=1)
    single, _ = _env("classic", "ft06/fifo")
    obs = venv.reset()
    sp = venv.observation_space
    for k in sp.keys():
        assert obs[k].shape == (B,) + tuple(sp[k].shape) and obs[k].dtype == sp[k].dtype, k
    rng = np.random.default_rng(0)
    episodes = 0
    for step in range(400):
        acts = rng.integers(0, 2, B)
        venv.step_async(acts)
        obs, rew, dones, infos = venv.step_wait()
        o1, r1, term, trunc, info1 = single.step(int(acts[3]))
        assert rew.dtype == np.float32 and np.float32(r1) == rew[3] and dones[3] == (term or trunc)
        if dones[3]:
            episodes += 1
            t = infos[3]["terminal_observation"]
            assert np.allclose(t["current_transition"], 1.0) and infos[3]["makespan"] == info1["makespan"]
            assert infos[3]["TimeLimit.truncated"] is False and infos[3]["episode"]["l"] > 36
            for k in sp.keys():
                assert np.array_equal(np.asarray(o1[k]).astype(t[k].dtype).reshape(t[k].shape), t[k]), k
            o1, _ = single.reset()
        for k in sp.keys():
            assert np.array_equal(np.asarray(o1[k]).astype(obs[k].dtype).reshape(obs[k][3].shape), obs[k][3]), (step, k)
    assert episodes >= 3
    assert venv.env_is_wrapped(object) == [False] * B and len(venv.get_attr("num_envs")) == B
    venv.close(); single.close()


def test_gym_surface_of_the_single_env():
    """env.py:347-358, :404-411, :533-536 of the reference: factories passed as classes / instances / names,
    observation_space, history, and no device allocation on reset when the instance is unchanged."""
    from jobshoplab_b200 import Config, JobShopLabEnv
    from jobshoplab_b200 import compiler as C
    from jobshoplab_b200.exceptions import ConfigurationError

    sc = G.Scenario("classic", "ft06/random0")
    li = C.compile_spec_text(sc.source_text)

    class BinaryActionObservationFactory:  # stands for the reference's class of that name
        pass

    class Mine:
        pass

    env = JobShopLabEnv(config=Config(), compiler=li, observation_factory=BinaryActionObservationFactory,
                        middleware="EventBasedBinaryActionMiddleware", seed=3)
    with pytest.raises(ConfigurationError):
        JobShopLabEnv(config=Config(), compiler=li, observation_factory=Mine)
    sp = env.observation_space
    assert set(sp.keys()) == {"job_running", "job_executed_on_machine", "job_progression", "machine_running",
                              "machine_progression", "available_jobs", "current_time", "current_transition"}
    assert sp["job_executed_on_machine"].shape == (6, 6) and sp["current_transition"].dtype == np.float32
    assert env.action_space.contains(1) and not env.action_space.contains(2)
    b0 = env._batch
    for ep in range(2):
        assert env.history == tuple()
        n_noop = 0
        for i, a in enumerate(sc.actions.tolist()):
            cand = env.state.possible_transitions[0]
            obs, r, term, trunc, info = env.step(int(a))
            assert r == float(sc.reward[i]) and info["no_op"] == (a == 0)
            n_noop += a == 0
            h = env.history[-1]
            assert len(env.history) == i + 1 and (len(h.action.transitions) == 0) == (a == 0)
            if a == 1:
                assert h.action.transitions[0].component_id == cand.component_id and h.action.transitions[0].job_id == cand.job_id
            assert info["no_op_count"] == n_noop == sum(len(h.action.transitions) == 0 for h in env.history)
        assert term and info["makespan"] == sc.meta["makespan"]
        env.reset()
        assert env._batch is b0  # same instance: the device batch is reused, reset is one kernel
    env.close()


@pytest.mark.gpu
def test_host_read_back_paths():
    """Batch.fetch_host (one env: single copy for a one-env batch, two otherwise) and Batch.fetch_host_all (every env,
    one copy) return the bytes the device tensors hold; host_observation builds the reference's dict from them."""
    import torch

    from jobshoplab_b200 import engine
    from jobshoplab_b200.middleware import host_observation

    sc = G.Scenario("transport", "ft10-t/random0")
    for B in (1, 5):
        b = engine.Batch(sc.lowered, B, **sc.cfg)
        b.reset()
        rng = np.random.default_rng(B)
        for i in range(12):
            b.step(torch.from_numpy(rng.integers(0, 2, B).astype(np.uint8)).cuda())
        rec_dev, obs_dev = b.record.cpu().numpy(), b.out.obs.cpu().numpy()
        rec_all, obs_all = b.fetch_host_all()
        assert np.array_equal(rec_all, rec_dev) and np.array_equal(obs_all, obs_dev)
        for idx in range(B):
            rec, obs = b.fetch_host(idx)
            assert np.array_equal(rec, rec_dev[idx]) and np.array_equal(obs, obs_dev[idx]), (B, idx)
            d = host_observation(b, idx)
            views = b.unpack_obs(b.out.obs)
            assert d["job_progression"] == tuple(int(x) for x in views["job_progression"][idx].cpu())
            assert d["job_executed_on_machine"] == tuple(tuple(bool(x) for x in row) for row in views["job_executed_on_machine"][idx].cpu())
            assert np.array_equal(d["current_transition"], views["current_transition"][idx].cpu().numpy())
        batched = b.unpack_obs_host(obs_all)
        for k, v in b.unpack_obs(b.out.obs).items():
            assert np.array_equal(np.asarray(batched[k]), v.cpu().numpy()), k
        b.close()
