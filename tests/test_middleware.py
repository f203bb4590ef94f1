"""Seam 2 of SURVEY 8(b): the CUDA engine plugged into the UNMODIFIED reference env through its own middleware slot.

    JobShopLabEnv(config, compiler=..., middleware=CudaEventBasedBinaryActionMiddleware)      # reference env

runs next to the stock env (``EventBasedBinaryActionMiddleware``) on the same instance, seed and action sequence;
every step must give the same observation, reward, flags, info, candidate list and -- compared as the canonical
digest AND as the reference's own dataclasses -- the same ``State``.  Needs the reference (oracle/_ref, made by
oracle/make_ref.py; /root/reference in the build container) and a GPU.
"""
import random
import sys
from dataclasses import replace
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))


@pytest.fixture(scope="module")
def rh():
    from oracle import ref_bench
    if not ref_bench.available():
        pytest.skip("the reference is not installed (python oracle/make_ref.py)")
    import ref_harness
    return ref_harness


def _envs(rh, kind, rel, seed, **cfg_over):
    from jobshoplab import JobShopLabEnv
    from jobshoplab.compiler import Compiler
    from jobshoplab.compiler.repos import DslRepository, SpecRepository
    from jobshoplab_b200.middleware import CudaEventBasedBinaryActionMiddleware

    path = rh.REF_ROOT / "data" / "jssp_instances" / rel
    out = []
    for mw in (None, CudaEventBasedBinaryActionMiddleware):
        cfg = rh.make_config(**cfg_over)
        repo = (SpecRepository if kind == "spec" else DslRepository)(path, loglevel="warning", config=cfg)
        out.append(JobShopLabEnv(config=cfg, compiler=Compiler(cfg, loglevel="warning", repo=repo), middleware=mw, seed=seed))
    return out


def _obs_equal(a, b):
    assert set(a) == set(b)
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and np.array_equal(x, y), (k, x, y)


def _neutral(state):
    """State with the one field the digest does not carry (TransportLocation.progress, a TODO upstream:
    handler.py:791) set to 0."""
    ts = tuple(replace(t, location=type(t.location)(progress=0, location=t.location.location)) for t in state.transports)
    return replace(state, transports=ts)


def _compare(rh, stock, cuda, tag):
    ix = rh.Indexer(stock.instance)
    assert np.array_equal(rh.digest(ix, stock.state), rh.digest(rh.Indexer(cuda.instance), cuda.state)), (tag, "digest")
    assert stock.state.possible_transitions == cuda.state.possible_transitions, (tag, "possible transitions")
    assert _neutral(stock.state.state) == _neutral(cuda.state.state), (tag, "State dataclasses")


CASES = [("spec", "spec_files/ft06", {}), ("dsl", "transport/ft10.yaml", {}),
         ("dsl", "transport/la01.yaml", {"allow_early_transport": False}),
         ("dsl", "dsl/instance_3x3_full.yaml", {}), ("spec", "spec_files/la16", {"truncation_active": True, "truncation_joker": 2})]


@pytest.mark.parametrize("kind,rel,over", CASES)
def test_reference_env_with_cuda_middleware_equals_stock(rh, kind, rel, over):
    for seed in (3, 11):
        stock, cuda = _envs(rh, kind, rel, seed, **over)
        assert type(cuda.state_simulator).__name__ == "CudaEventBasedBinaryActionMiddleware"
        _obs_equal(stock.current_observation[0], cuda.current_observation[0])
        _compare(rh, stock, cuda, (rel, seed, "reset"))
        rng = random.Random(seed)
        steps = 0
        while not stock.done:
            a = rng.randint(0, 1) if steps % 7 else 1
            so, sr, st, str_, si = stock.step(a)
            co, cr, ct, ctr, ci = cuda.step(a)
            tag = (rel, seed, steps, a)
            _obs_equal(so, co)
            assert sr == cr and st == ct and str_ == ctr and si == ci, (tag, sr, cr, si, ci)
            assert stock.state.message == cuda.state.message or not cuda.state.success, tag
            _compare(rh, stock, cuda, tag)
            steps += 1
        assert cuda.done and len(stock.history) == len(cuda.history)
        # a second episode through env.reset(): the device batch is reused when the instance is unchanged
        b0 = cuda.state_simulator._batch
        stock.reset(); cuda.reset()
        _compare(rh, stock, cuda, (rel, seed, "second reset"))
        if kind == "spec":  # (the env made a new middleware object; the device batch is the cached one)
            assert cuda.state_simulator._batch is b0
        with pytest.raises(Exception) as e1:
            stock.step(5)
        with pytest.raises(Exception) as e2:
            cuda.step(5)
        assert type(e1.value).__name__ == type(e2.value).__name__ == "ActionOutOfActionSpace"
        cuda.state_simulator.close()


def test_dispatch_rules_run_on_the_plugged_env(rh):
    """utils/dipatching_benchmark.py:17-67 drives `env.state` / `env.instance` of the reference env: the rules read the
    rebuilt State (prebuffer stores, operation durations) and reach the reference's makespans."""
    for rule, want in (("fifo", 1262), ("spt", 1074), ("mwkr", 1262)):
        _, cuda = _envs(rh, "spec", "spec_files/ft10", 0)
        while not cuda.done:
            cuda.step(rh.rule_action(cuda, rule))
        assert cuda.state.state.time.time == want, rule
        cuda.state_simulator.close()


def test_user_observation_factory_on_the_plugged_env(rh):
    """a user-defined ObservationFactory (reference ABC, observations.py:31-76) works unchanged -- it
    is called on the rebuilt StateMachineResult."""
    from jobshoplab import JobShopLabEnv
    from jobshoplab.compiler import Compiler
    from jobshoplab.compiler.repos import SpecRepository
    from jobshoplab.env.factories.observations import ObservationFactory
    from jobshoplab_b200.middleware import CudaEventBasedBinaryActionMiddleware
    import gymnasium as gym

    class RemainingOps(ObservationFactory):
        def __init__(self, loglevel, config, instance, *a, **k):
            super().__init__(loglevel, config, instance)
            self.observation_space = gym.spaces.Box(low=0, high=100, shape=(1,), dtype=np.int32)

        def __repr__(self):
            return "RemainingOps"

        def make(self, state_result, *a, **k):
            return np.asarray([sum(o.operation_state_state.value != "Done" for j in state_result.state.jobs for o in j.operations)])

    envs = []
    for mw in (None, CudaEventBasedBinaryActionMiddleware):
        cfg = rh.make_config()
        repo = SpecRepository(rh.REF_ROOT / "data" / "jssp_instances" / "spec_files" / "ft06", loglevel="warning", config=cfg)
        envs.append(JobShopLabEnv(config=cfg, compiler=Compiler(cfg, loglevel="warning", repo=repo), middleware=mw,
                                  observation_factory=RemainingOps))
    stock, cuda = envs
    while not stock.done:
        so, sr, *_ = stock.step(1)
        co, cr, *_ = cuda.step(1)
        assert np.array_equal(so, co) and sr == cr
    assert so[0] == 0 and cuda.done
