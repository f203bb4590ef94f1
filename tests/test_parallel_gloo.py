"""World-size-2 (gloo, CPU) test of the multi-GPU host logic: contiguous env sharding, shard-invariant
instance generation and the final statistics all-reduce equal the single-process result.  The per-rank
"rollout" is the CPU oracle here (checker role); on the GPU box the same code path runs with NCCL."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _episodes(first, count):
    import bench
    from jobshoplab_b200 import abi
    from oracle import oracle_py as op
    import copy

    tmpl = bench.template_instance()
    om, od = bench.make_instances(count, first)
    lb, mat = bench.taillard_bounds(om, od)
    insts = []
    for i in range(count):
        li = copy.copy(tmpl)
        li.op_machine, li.op_duration = om[i].copy(), od[i].copy()
        li.lower_bound, li.max_allowed_time = int(lb[i]), int(mat[i])
        insts.append(li)
    r = op.rollout_batch(insts, count, abi.POLICY_RANDOM, seed=0, n_threads=2, env_base=first)
    return {k: torch.from_numpy(v) for k, v in r.items() if k != "total_steps"}


def _worker(rank, world, total, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jobshoplab_b200 import parallel

    first, count = parallel.shard_range(total, rank, world)
    r = _episodes(first, count)
    st = parallel.reduce_stats(r["makespan"], r["n_steps"], r["ret"], r["status"])
    if rank == 0:
        q.put((st.count, st.terminated, st.steps, st.makespan_sum, st.makespan_min, st.makespan_max, st.return_sum,
               st.histogram.tolist()))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from jobshoplab_b200 import parallel

    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b


def test_two_rank_reduction_equals_single_process():
    from jobshoplab_b200 import parallel

    total = 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, total, 29517, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    r = _episodes(0, total)
    st = parallel.reduce_stats(r["makespan"], r["n_steps"], r["ret"], r["status"])
    want = (st.count, st.terminated, st.steps, st.makespan_sum, st.makespan_min, st.makespan_max, st.return_sum,
            st.histogram.tolist())
    assert got[:3] == want[:3] and got[4:6] == want[4:6] and got[7] == want[7]
    assert abs(got[3] - want[3]) < 1e-9 and abs(got[6] - want[6]) < 1e-9
    assert st.terminated == total
