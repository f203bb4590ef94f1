"""What env.render() draws: the reference dashboard's data model,
DashboardDataMapper.map_states_to_gant_data(env.history) (rendering/gant_dashboard.py:414-429), rebuilt from the
engine's state digest and its transport-task log.  Golden vectors: tests/golden/gantt.json, recorded from the
unmodified reference by oracle/gen_gantt_golden.py half way through and at the end of stored action tapes.
CPU: the engine source compiled for the host; GPU (-m gpu): the CUDA engine through the C ABI."""
import json
from pathlib import Path

import pytest

import _golden as G
from jobshoplab_b200.state_view import gantt_data, parse_digest

GOLD = json.loads((Path(__file__).parent / "golden" / "gantt.json").read_text())


def _norm(d):
    return {k: [dict(x) for x in v] for k, v in d.items()}


def _check(key, snapshots):
    """snapshots(steps) -> gantt dict after that many steps of the scenario's tape"""
    rec = GOLD[key]
    for which in ("half", "final"):
        if which not in rec:
            continue
        got = _norm(snapshots(rec[which]["steps"]))
        want = rec[which]["data"]
        assert got["schedules"] == want["schedules"], (key, which, "schedules")
        assert got["transports"] == want["transports"], (key, which, "transports")
        assert got["buffer"] == want["buffer"] == []


@pytest.mark.parametrize("key", sorted(GOLD))
def test_gantt_data_engine_source(key):
    from hostsim.hostsim_py import HostSimEnv

    group, name = key.split(":")
    sc = G.Scenario(group, name)

    def snapshots(steps):
        e = HostSimEnv(sc.lowered, **sc.cfg)
        for i in range(steps):
            e.step(int(sc.actions[i]))
        return gantt_data(sc.lowered, parse_digest(sc.lowered, e.digest()), e.leg_log())

    _check(key, snapshots)


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLD))
def test_gantt_data_gpu(key):
    import torch
    from jobshoplab_b200 import engine

    group, name = key.split(":")
    sc = G.Scenario(group, name)

    def snapshots(steps):
        b = engine.Batch(sc.lowered, 2, record_ops=True, **sc.cfg)
        b.reset()
        a = torch.zeros(2, dtype=torch.uint8, device="cuda")
        for i in range(steps):
            a.fill_(int(sc.actions[i]))
            b.step(a)
        out = gantt_data(sc.lowered, parse_digest(sc.lowered, b.get_state(1)), b.get_leg_log(1))
        b.close()
        return out

    _check(key, snapshots)
