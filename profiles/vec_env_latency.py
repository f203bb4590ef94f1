"""Host-loop cost of vec_env.SB3VecEnv (what a stable-baselines3 PPO loop pays per step): env-steps/s over 200 steps."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import _golden as G
from jobshoplab_b200 import vec_env
mods = {"shipped": vec_env}
try:
    sys.path.insert(0, str(ROOT / "profiles"))
    import _old_vec_env
    mods["previous (per-field copies)"] = _old_vec_env
except ImportError:
    pass
sc = G.Scenario("classic", "ft10/random0")
for B in (256, 4096):
    for name, mod in mods.items():
        from jobshoplab_b200.env import BatchedJobShopLabEnv
        venv = mod.SB3VecEnv(venv=BatchedJobShopLabEnv(instances=sc.lowered, num_envs=B))
        rng = np.random.default_rng(0)
        venv.reset()
        acts = rng.integers(0, 2, (220, B))
        for i in range(20): venv.step(acts[i])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(20, 220): venv.step(acts[i])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"SB3VecEnv {name}: B={B}: {1e3 * dt / 200:.3f} ms per vector step, {200 * B / dt / 1e6:.3f} M env-steps/s")
        venv.close()
