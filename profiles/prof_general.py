"""Workload for ncu captures of the non-classic kernels: 3 fused rollouts under the random policy over replicas of
ta41-t (30x20 with travel times: the "simple" level) or, with `fifo`, la01-t with FIFO buffers of capacity 5
(BASELINE config 4: the "queues" level) or, with `stoch`, the stochastic ft20-t variant (setup times, gamma breakdowns,
Poisson travel: BASELINE config 5, the general stochastic kernel) under the always-accept policy.

    python profiles/prof_general.py [envs] [fifo|stoch]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _golden as G  # noqa: E402
from jobshoplab_b200 import compiler, engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
if len(sys.argv) > 2 and sys.argv[2] == "fifo":
    sc = G.Scenario("buffers", "la01-t/fifo5/random0")
    batch = engine.Batch(sc.lowered, B, **sc.cfg)
elif len(sys.argv) > 2 and sys.argv[2] == "stoch":
    batch = engine.Batch(G.KsFixture("ft20-t_features").lowered, B, seed=5)
else:
    li = compiler.compile_dsl_text(G.Scenario("transport", "ta41-t/fifo").source_text)
    batch = engine.Batch(li, B)
for _ in range(3):
    r = batch.rollout("fifo" if sys.argv[2:3] == ["stoch"] else "random", reset_first=True, writeback=False)
torch.cuda.synchronize()
print("rollout env-steps per launch:", int(r.n_steps.sum()))
