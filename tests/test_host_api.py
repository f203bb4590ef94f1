"""Host-side mirror of the reference API that needs no GPU: config loading (load_config.py:15-46), the
state views rebuilt from a canonical digest, the C-ABI library's exported symbols, and the loud failure
when no CUDA device is present."""
import ctypes
import re
import sys
from pathlib import Path

import pytest

import _golden as G
from jobshoplab_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


def test_load_config_groups(tmp_path):
    from jobshoplab_b200 import load_config

    p = tmp_path / "cfg.yaml"
    p.write_text("""
title: t
compiler: {repo: SpecRepository, spec_repository: {dir: data/x/ft06}}
env: {observation_factory: BinaryActionObservationFactory}
state_machine: {allow_early_transport: false}
middleware: {event_based_binary_action_middleware: {truncation_joker: 3, truncation_active: true}}
reward_factory: {binary_action_jssp_reward: {sparse_bias: 1, dense_bias: 0.001, truncation_bias: -1}}
render_backend: {render_in_dashboard: {port: 8050}}
""")
    c = load_config(p, overrides=["state_machine.allow_early_transport=true", "compiler.spec_repository.dir=zz"])
    assert c.compiler.repo == "SpecRepository" and c.compiler.spec_repository.dir == "zz"
    assert c.state_machine.allow_early_transport is True
    assert c.middleware.event_based_binary_action_middleware.truncation_joker == 3
    assert isinstance(c.reward_factory.binary_action_jssp_reward.sparse_bias, float)
    assert "render_backend" in c.extra


def test_state_view_from_digest():
    from jobshoplab_b200.state_view import parse_digest
    from oracle import oracle_py as op

    sc = G.Scenario("transport", "ft06-t/fifo")
    env = op.OracleEnv(sc.lowered)
    v = parse_digest(sc.lowered, env.digest())
    assert v.state.time.time == 0 and len(v.state.jobs) == 6 and len(v.state.transports) == 6
    assert len(v.possible_transitions) == 36 and v.possible_transitions[0].component_id == "t-0"
    assert v.possible_transitions[0].job_id == "j-0" and v.possible_transitions[0].new_state == "Working"
    assert v.state.jobs[0].location == "b-24" and v.state.transports[3].location == "m-3"
    for a in sc.actions[:40]:
        env.step(int(a))
    v = parse_digest(sc.lowered, env.digest())
    assert v.state.time.time == env.time
    assert sum(len(m.prebuffer) + len(m.buffer) + len(m.postbuffer) for m in v.state.machines) + \
        sum(t.carried is not None for t in v.state.transports) + sum(len(s) for _, s in v.state.buffers) == 6


def test_abi_library_exports_every_declared_symbol():
    from jobshoplab_b200 import engine

    if not (engine.LIB_PATH.exists() and engine.LOG_LIB_PATH.exists()):
        from jobshoplab_b200 import build
        build.build()
    header = (ROOT / "include" / "jobshoplab_b200.h").read_text()
    declared = set(re.findall(r"\b(jsl_[a-z_]+)\s*\(", header))
    assert declared >= set(engine.ABI_SYMBOLS) and len(declared) >= 15
    for path in (engine.LIB_PATH, engine.LOG_LIB_PATH):  # throughput build and logging build: one ABI
        lib = ctypes.CDLL(str(path))
        for sym in declared:
            getattr(lib, sym)
        assert lib.jsl_abi_version() == abi.ABI_VERSION == 2
    # constants the Python mirror restates
    defines = dict(re.findall(r"#define\s+(JSL_[A-Z_0-9]+)\s+(\d+)u?\b", header))
    assert int(defines["JSL_STEP_AUTO_RESET"]) == abi.STEP_AUTO_RESET
    assert int(defines["JSL_OBS_STATE"]) == abi.OBS_STATE
    import ctypes as C
    assert C.sizeof(abi.StepOut) == 9 * 8 + 8 and abi.STEP_RECORD_BYTES == 32


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jobshoplab_b200 import engine
    from jobshoplab_b200.exceptions import EngineUnavailable

    sc = G.Scenario("classic", "ft06/fifo")
    with pytest.raises(EngineUnavailable):
        engine.Batch(sc.lowered, 4)
    for p in (ROOT / "jobshoplab_b200").glob("*.py"):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "hostsim" not in text, p


def test_time_fraction_shortcut_is_exact():
    """jsl_engine.cuh time_fraction(): float32(t * (1/m)) == float32(t / m) for every integer pair the engine feeds it
    (m < 2^22): exhaustive for small m, random for large."""
    import numpy as np

    for m in range(1, 1500):
        t = np.arange(0, 3 * m + 5, dtype=np.float64)
        assert np.array_equal((t / m).astype(np.float32), (t * (1.0 / m)).astype(np.float32)), m
    rng = np.random.default_rng(1)
    m = rng.integers(1, 1 << 22, size=4_000_000).astype(np.float64)
    t = np.floor(rng.random(4_000_000) * m * 1.5)
    assert np.array_equal((t / m).astype(np.float32), (t * (1.0 / m)).astype(np.float32))


def test_host_unpack_matches_tensor_unpack():
    import numpy as np
    """Batch.unpack_obs_host (numpy views of one env's bytes, the single-env read-back) names the same bytes as
    Batch.unpack_obs (torch views of the [B, bytes] record) for every packed observation kind."""
    import types

    import torch

    from jobshoplab_b200 import abi, engine

    rng = np.random.default_rng(5)
    J, M, N = 7, 5, 31
    for kind, nbytes in ((abi.OBS_BINARY_ACTION, (16 + 4 * J + 4 * M + 2 * J + M + J * M + 15) & ~15),
                         (abi.OBS_OPERATION_ARRAY, 16 + 4 * N + 4 * J), (abi.OBS_TASSEL, 28 * J)):
        stub = types.SimpleNamespace(shape=types.SimpleNamespace(n_jobs=J, n_machines=M, n_ops=N),
                                     cfg=types.SimpleNamespace(obs_kind=kind), torch=torch, out=None)
        raw = rng.integers(0, 256, nbytes, dtype=np.uint8)
        host = engine.Batch.unpack_obs_host(stub, raw)
        dev = engine.Batch.unpack_obs(stub, torch.from_numpy(raw.copy())[None])
        assert set(host) == set(dev)
        for k in host:
            want = dev[k][0].numpy()
            assert host[k].dtype == want.dtype and host[k].shape == want.shape, (kind, k)
            assert host[k].tobytes() == want.tobytes(), (kind, k)
        # all envs at once (Batch.fetch_host_all): same entries with the leading batch axis
        raw2 = rng.integers(0, 256, (5, nbytes), dtype=np.uint8)
        host2 = engine.Batch.unpack_obs_host(stub, raw2)
        dev2 = engine.Batch.unpack_obs(stub, torch.from_numpy(raw2.copy()))
        for k in host2:
            want = dev2[k].numpy()
            assert host2[k].dtype == want.dtype and host2[k].shape == want.shape, (kind, k, host2[k].shape, want.shape)
            assert np.array(host2[k]).tobytes() == np.ascontiguousarray(want).tobytes(), (kind, k)
