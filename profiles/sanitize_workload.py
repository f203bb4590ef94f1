"""Tiny workload for compute-sanitizer (one tool per gpurun call): classic, transport with LIFO capacity-2
buffers (general path incl. TimeDependency / BufferFull), stochastic table mode; rollout + step mode."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _golden as G  # noqa: E402
from jobshoplab_b200 import engine  # noqa: E402

for group, name in (("classic", "ft06/random0"), ("buffers", "la01-t/lifo2/noearly/random0"),
                    ("features", "full_feature_3x3_instance/random0"), ("solutions", "swv16/solution")):
    sc = G.Scenario(group, name)
    B = 6
    batch = engine.Batch(sc.lowered, B, record_ops=True, **sc.cfg)
    if not sc.lowered.is_deterministic:
        batch.set_start_seeds(torch.from_numpy(np.tile(sc.start_seeds, (B, 1)).astype(np.int32)).cuda())
    batch.reset()
    n = min(len(sc.actions), 60)
    acts = torch.from_numpy(np.tile(sc.actions[:n], (B, 1)).astype(np.uint8)).cuda()
    for i in range(n):
        batch.step(acts[:, i].contiguous())
    d = batch.get_state(B - 1)
    batch.reset()
    r = batch.rollout("random", max_steps=300)
    torch.cuda.synchronize()
    print(group, name, "steps", n, "digest len", len(d), "rollout steps", r.n_steps.cpu().tolist()[:3], "status", r.status.cpu().tolist()[:3])
    batch.close()
    # the throughput build as well, with the raw-state observation (image-sized observation copies) and, for the
    # classic instance, uploaded / device-randomised variant tables
    fast = engine.Batch(sc.lowered, B, obs_kind=4, **{k: v for k, v in sc.cfg.items() if k != "obs_kind"})
    fast.reset()
    for i in range(min(n, 30)):
        fast.step(acts[:, i].contiguous())
    fast.rollout("fifo", max_steps=200)
    if sc.lowered.is_deterministic and group == "classic":
        om = np.tile(sc.lowered.op_machine, (B, 1)).astype(np.int32); od = np.tile(sc.lowered.op_duration, (B, 1)).astype(np.int32)
        lb = np.full(B, sc.lowered.lower_bound, np.int32); mat = np.full(B, sc.lowered.max_allowed_time, np.int32)
        vb = engine.Batch(sc.lowered, B, variants=(om, od, lb, mat))
        vb.load_variants(om, od, lb, mat)
        vb.randomize(seed=3)
        vb.rollout("random", reset_first=True)
        vb.check()
        vb.close()
    torch.cuda.synchronize()
    fast.close()
print("sanitize workload done")
