import time, sys, cProfile, pstats
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import _golden as G
from jobshoplab_b200.env import JobShopLabEnv
for name in ("ft06/random0",):
    sc = G.Scenario("classic", name)
    env = JobShopLabEnv(compiler=sc.lowered, seed=1)
    rng = np.random.default_rng(0)
    for rep in range(3):
        env.reset(); n = 0; t0 = time.perf_counter()
        while not env.done:
            env.step(int(rng.integers(0, 2))); n += 1
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(name, "steps", n, "steps/s", n / dt, "us/step", 1e6 * dt / n)
    pr = cProfile.Profile(); env.reset(); pr.enable()
    while not env.done: env.step(int(rng.integers(0, 2)))
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
