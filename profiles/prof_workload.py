"""Small fixed workload for ncu captures: ta41-shape batch, 3 rollout launches then `steps` step launches.

    python profiles/prof_workload.py [envs] [steps]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from jobshoplab_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 12
tmpl = bench.template_instance()
om, od = bench.make_instances(B)
lb, mat = bench.taillard_bounds(om, od)
batch = engine.Batch(tmpl, B, seed=0, variants=(om, od, lb, mat), obs_kind=1)
for _ in range(3):
    r = batch.rollout("random", reset_first=True, writeback=False)
torch.cuda.synchronize()
print("rollout env-steps per launch:", int(r.n_steps.sum()))
batch.reset()
acts = torch.randint(0, 2, (STEPS, B), dtype=torch.uint8, device="cuda")
for i in range(STEPS):
    batch.step(acts[i])
torch.cuda.synchronize()
print("step launches done, envs:", B)
