"""User-defined observation / reward hooks (the reference's ABCs: observations.py:31-76, rewards.py:11-58).

Three routes: (1) torch-side functions over the raw state record (observation kind "StateObservation" =
JSL_OBS_STATE, named by jsl_batch_state_layout); (2) a device-side reward hook compiled into the library
(-DJSL_USER_HEADER, csrc/examples/jsl_user_example.cuh -> libjsl_b200_user.so); (3) the reference's own factory
objects on the reference env with the CUDA middleware (tests/test_middleware.py).
"""
import numpy as np
import pytest

import _golden as G
from jobshoplab_b200 import abi

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_state_record_and_torch_side_factories():
    from jobshoplab_b200 import BatchedJobShopLabEnv, Config
    from jobshoplab_b200 import compiler as C
    from oracle import oracle_py as op

    sc = G.Scenario("classic", "ft10/random0")
    li = C.compile_spec_text(sc.source_text)
    cfg = Config()
    cfg.env.observation_factory = "StateObservation"
    J, M = li.n_jobs, li.n_machines

    def remaining_work(views, out):  # a user observation: per job, how many operations are not DONE yet
        return (M - views["j_k"][:, :J].to(torch.int32))

    def busy_reward(views, out):     # a user reward: stock reward + fraction of WORKING machines
        return out.reward + (views["m_state"][:, :M] == abi.M_WORKING).double().mean(dim=1)

    B = 5
    venv = BatchedJobShopLabEnv(cfg, num_envs=B, compiler=li, observation_fn=remaining_work, reward_fn=busy_reward)
    lay = venv.batch.state_layout()
    assert {"time", "j_k", "j_proc", "j_loc", "m_state", "m_occ", "t_state", "ret", "max_allowed_time"} <= set(lay)
    assert max(o + dt.itemsize * n for o, dt, n in lay.values()) <= venv.batch.shape.state_bytes
    obs, _ = venv.reset()
    assert obs.shape == (B, J) and (obs == M).all()
    ora = op.OracleEnv(li)
    acts = torch.zeros(B, dtype=torch.uint8, device="cuda")
    for i, a in enumerate(sc.actions.tolist()):
        acts.fill_(int(a))
        obs, rew, term, trunc, out = venv.step(acts)
        ora.step(int(a))
        views = venv.batch.unpack_state()
        assert int(views["time"][2]) == ora.time and int(views["n_steps"][0]) == i + 1
        d = ora.digest()  # per job: n_done, running, loc (+ 3 per op) after the 5-int header
        n_done = [int(d[5 + j * (3 + 3 * M)]) for j in range(J)]
        assert obs[1].tolist() == [M - k for k in n_done]
        m_working = sum(int(x) == abi.M_WORKING for x in views["m_state"][0, :M].tolist())
        assert abs(float(rew[0]) - (ora.reward + m_working / M)) < 1e-12
    assert bool(term.all())
    venv.close()


def test_device_side_reward_hook_library(monkeypatch):
    """libjsl_b200_user.so = the same engine compiled with csrc/examples/jsl_user_example.cuh: reward' = reward -
    0.001 * (#machines not WORKING) / M, evaluated in the step and rollout kernels."""
    from jobshoplab_b200 import build, engine
    if not build.USER_EXAMPLE_LIB.exists():
        pytest.skip("libjsl_b200_user.so not built (jobshoplab_b200.build.build_user_example)")
    sc = G.Scenario("classic", "ft06/random0")
    stock = engine.Batch(sc.lowered, 3, obs_kind=abi.OBS_STATE)
    monkeypatch.setattr(engine, "LIB_PATH", build.USER_EXAMPLE_LIB)
    monkeypatch.setitem(engine._libs, False, None)
    user = engine.Batch(sc.lowered, 3, obs_kind=abi.OBS_STATE)
    assert user.L is not stock.L
    stock.reset(); user.reset()
    acts = torch.zeros(3, dtype=torch.uint8, device="cuda")
    M = sc.lowered.n_machines
    total = 0.0
    for a in sc.actions.tolist():
        acts.fill_(int(a))
        so, uo = stock.step(acts), user.step(acts)
        idle = int((user.unpack_state()["m_state"][0, :M] != abi.M_WORKING).sum())
        want = float(so.reward[0]) - 0.001 * idle / M
        assert float(uo.reward[0]) == want
        total += want
        assert torch.equal(so.obs[:, :48], uo.obs[:, :48])  # same trajectory (scalars up to `status`)
    r = user.rollout("tape", tape=torch.from_numpy(np.tile(sc.actions, (3, 1)).astype(np.uint8)).cuda(), reset_first=True)
    assert abs(float(r.ret[1]) - total) < 1e-12 and int(r.makespan[1]) == sc.meta["makespan"]
    monkeypatch.setitem(engine._libs, False, None)
    stock.close(); user.close()
