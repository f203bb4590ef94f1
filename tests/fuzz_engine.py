"""Differential fuzzing of the CUDA engine SOURCE (CPU build, tests/hostsim) against the oracle on random
small instances: random shapes, travel matrices with zeros, FIFO/LIFO/DUMMY/FLEX buffers with random
capacities, setup-time matrices, deterministic outages, custom standalone buffers, early transport on/off,
truncation.  Every step compares digest, reward, flags and observation; escaping statuses must agree.

    python tests/fuzz_engine.py [n_instances] [seed]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from hostsim.hostsim_py import HostSimEnv  # noqa: E402
from jobshoplab_b200 import lowered as L  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def random_instance(rng, simple=None, jrange=(1, 8)) -> tuple[L.LoweredInstance, dict]:
    """simple: None = anything; 1 = a 'simple' instance (FLEX unbounded buffers, no setup / outages: specialisation
    level 1 of the engine); 2 = 'classic' (simple with zero travel times: level 2, the fast paths); 3 = 'queues' (any
    buffer types and capacities but no setup times / outages: level 3); 4 = 'features' (setup times / outages allowed,
    FLEX unbounded buffers: level 4)."""
    J, M = int(rng.integers(*jrange)), int(rng.integers(1, 6))
    T = int(rng.integers(1, min(J, 127) + 2))
    S = 2
    nt = int(rng.integers(1, 4))
    job_op_start = [0]
    op_machine, op_tool, op_dur = [], [], []
    for _ in range(J):
        k = int(rng.integers(1, M + 2))
        for _ in range(k):
            op_machine.append(int(rng.integers(0, M)))
            op_tool.append(int(rng.integers(0, nt)))
            op_dur.append(int(rng.integers(0, 12)))
        job_op_start.append(len(op_machine))
    def btypes(n):
        if rng.random() < 0.4:
            return [L.BUF_FLEX] * n
        return [int(rng.choice([L.BUF_FIFO, L.BUF_LIFO, L.BUF_DUMMY, L.BUF_FLEX])) for _ in range(n)]
    def caps(n):
        if rng.random() < 0.5:
            return [L.CAP_INF] * n
        if J > 32:  # wide instances: capacities that let an episode get somewhere
            return [int(rng.choice([max(4, J // 3), J, L.CAP_INF, L.CAP_INF])) for _ in range(n)]
        return [int(rng.choice([1, 2, 3, J, L.CAP_INF])) for _ in range(n)]
    NP = M + S
    if rng.random() < 0.3:
        travel = np.zeros((NP, NP), dtype=np.int32)
    else:
        travel = rng.integers(0, 6, size=(NP, NP)).astype(np.int32)
        travel[rng.random((NP, NP)) < 0.3] = 0
        np.fill_diagonal(travel, 0)
    setup = np.zeros((M, nt, nt), dtype=np.int32) if rng.random() < 0.5 else rng.integers(0, 5, size=(M, nt, nt)).astype(np.int32)
    m_out_start, m_of, m_od = [0], [], []
    if rng.random() < 0.35:
        for m in range(M):
            for _ in range(int(rng.integers(0, 3))):
                m_of.append(int(rng.integers(0, 40))); m_od.append(int(rng.integers(0, 6)))
            m_out_start.append(len(m_of))
    else:
        m_out_start = [0] * (M + 1)
    t_of, t_od = [], []
    if rng.random() < 0.3:
        for _ in range(int(rng.integers(1, 3))):
            t_of.append(int(rng.integers(0, 30))); t_od.append(int(rng.integers(0, 5)))
    if simple == 3:
        setup = np.zeros((M, nt, nt), dtype=np.int32)
        m_out_start, m_of, m_od, t_of, t_od = [0] * (M + 1), [], [], [], []
    elif simple == 4:
        btypes = lambda n: [L.BUF_FLEX] * n
        caps = lambda n: [L.CAP_INF] * n
    elif simple:
        btypes = lambda n: [L.BUF_FLEX] * n
        caps = lambda n: [L.CAP_INF] * n
        setup = np.zeros((M, nt, nt), dtype=np.int32)
        m_out_start, m_of, m_od, t_of, t_od = [0] * (M + 1), [], [], [], []
        if simple == 2:
            travel = np.zeros((NP, NP), dtype=np.int32)
        elif not travel.any():
            travel[0, NP - 1] = 2
    NB = 3 * M + T + S
    li = L.LoweredInstance(
        description="fuzz", n_jobs=J, n_machines=M, n_transports=T, n_sbuf=S, n_tools=nt,
        job_op_start=job_op_start, op_machine=op_machine, op_tool=op_tool, op_duration=op_dur,
        pre_type=btypes(M), pre_cap=caps(M), post_type=btypes(M), post_cap=caps(M),
        setup_time=setup.reshape(-1), default_tool=0,
        sbuf_type=[int(rng.choice([L.BUF_FLEX, L.BUF_FIFO, L.BUF_LIFO])) if (rng.random() < 0.3 and simple in (None, 3)) else L.BUF_FLEX, L.BUF_FLEX],
        sbuf_cap=[L.CAP_INF, L.CAP_INF if (rng.random() < 0.8 or simple in (1, 2, 4)) else max(1, J - 1)],
        sbuf_role=[L.ROLE_INPUT, L.ROLE_OUTPUT], buf_ref_id=np.arange(NB), travel_time=travel.reshape(-1),
        agv_init_place=[int(rng.integers(0, NP if rng.random() < 0.2 else M)) for _ in range(T)],
        job_init_buf=[3 * M + T] * J, start_time=int(rng.choice([0, 0, 3])),
        m_outage_start=m_out_start, m_outage_freq=m_of, m_outage_dur=m_od, t_outage_freq=t_of, t_outage_dur=t_od,
        lower_bound=1, max_allowed_time=max(1, sum(op_dur)) + 1, tools=[f"tl-{i}" for i in range(nt)])
    cfg = dict(allow_early_transport=bool(rng.random() < 0.6), truncation_active=bool(rng.random() < 0.3),
               truncation_joker=int(rng.integers(0, 4)))
    return li, cfg


def run_one(rng, max_steps=400, simple=None, jrange=(1, 8)):
    li, cfg = random_instance(rng, simple, jrange)
    e, o = HostSimEnv(li, **cfg), op.OracleEnv(li, **cfg)
    p = float(rng.choice([0.5, 0.8, 1.0]))
    steps = 0

    def check(tag):
        if e.status != o.status:
            raise AssertionError(f"{tag}: status {e.status} vs {o.status}")
        if tag == "reset" and o.status > 3:
            return  # the reference's constructor raised: there is no env to compare
        a, b = e.digest(), o.digest()
        if not np.array_equal(a, b):
            raise AssertionError(f"{tag}: status {e.status} vs {o.status}\n eng {a.tolist()}\n ora {b.tolist()}")

    check("reset")
    while steps < max_steps and o.status == 0 and e.status == 0:
        a = int(rng.random() < p)
        se, so = e.step(a), o.step(a)
        assert se == so, (steps, se, so)
        check(f"step {steps} a={a}")
        if so <= 3:
            assert e.reward == o.reward, (steps, e.reward, o.reward)
            assert e.terminated == o.terminated and e.truncated == o.truncated
            assert e.obs() == o.obs(), (steps, "obs")
        steps += 1
    return steps, o.status


def main(n=300, seed=0):
    rng = np.random.default_rng(seed)
    total, hist = 0, {}
    for i in range(n):
        sub = np.random.default_rng(rng.integers(0, 2**63))
        state = sub.bit_generator.state
        try:
            steps, st = run_one(sub, simple=(None, 3, 1, 2, 4)[i % 5])  # a fifth general, a fifth each specialised level
        except AssertionError as ex:
            print(f"FUZZ FAILURE at instance {i} (seed {seed}): {str(ex)[:3000]}")
            raise SystemExit(1)
        total += steps
        hist[st] = hist.get(st, 0) + 1
    print(f"fuzz ok: {n} instances, {total} steps, final statuses {dict(sorted(hist.items()))}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
