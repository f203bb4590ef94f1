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

# BatchedJobShopLabEnv(auto_reset=True): the reset fused into the step launch against step + flags -> host -> jsl_reset
from jobshoplab_b200 import abi
from jobshoplab_b200.env import BatchedJobShopLabEnv
for B in (4096, 65536, 1048576):
    for mode in ("fused (JSL_STEP_AUTO_RESET)", "step, host decision, masked reset"):
        env = BatchedJobShopLabEnv(instances=sc.lowered, num_envs=B, auto_reset=mode.startswith("fused"))
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        acts = (torch.rand((260, B), device="cuda", generator=g) < 0.7).to(torch.uint8)
        def one(i):
            obs, reward, term, trunc, out = env.step(acts[i])
            if not env.auto_reset:  # what the env did before the flag existed
                done = (out.terminated | out.truncated | (out.status >= abi.ST_STEP_FAILED).to(out.terminated.dtype)).contiguous()
                if bool(done.any()):
                    final = env.batch.record.clone()
                    env.batch.reset(done)
                    rec = env.batch.record
                    rec[:, 0:8] = final[:, 0:8]; rec[:, 12:16] = final[:, 12:16]; rec[:, 24:26] = final[:, 24:26]
        for i in range(60): one(i)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(60, 260): one(i)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"auto-reset {mode}: B={B}: {1e3 * dt / 200:.3f} ms per vector step, {200 * B / dt / 1e6:.1f} M env-steps/s")
        env.close()
