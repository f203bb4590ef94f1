"""More distribution fixtures for the stochastic path: reference makespan samples of the two KS
instances of oracle/gen_golden.py (`ks_sources`) under the spt rule and under uniform-random actions, next to the
always-accept samples already in tests/golden/ks.npz.  TEST INFRASTRUCTURE ONLY (runs the unmodified reference).

    python oracle/gen_ks_more.py          # ~10 min on 8 cores -> tests/golden/ks2.npz
"""
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))
import gen_golden as gg  # noqa: E402


def episode(args):
    import ref_harness as rh

    text, seed, policy = args
    env = rh.make_env("dslstr", text, seed=seed)
    n = 0
    while not env.done:
        a = rh.random_action(seed, 0, n) if policy == "random" else rh.rule_action(env, policy)
        env.step(a)
        n += 1
    return int(env.state.state.time.time), n, int(env.terminated)


def main():
    out, names = {}, []
    for name, (text, n) in gg.ks_sources().items():
        for policy in ("spt", "random"):
            t0 = time.time()
            with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
                res = pool.map(episode, [(text, s, policy) for s in range(n)], chunksize=8)
            mk = np.sort(np.asarray([r[0] for r in res], dtype=np.int32))
            key = f"{name}/{policy}"
            names.append(key)
            out[f"{key}|makespan"] = mk
            out[f"{key}|n_steps"] = np.asarray([r[1] for r in res], dtype=np.int32)
            out[f"{key}|terminated"] = np.asarray([r[2] for r in res], dtype=np.uint8)
            print(f"{key}: {n} reference episodes, makespan mean {mk.mean():.1f} std {mk.std():.1f}, "
                  f"terminated {int(sum(r[2] for r in res))}, {time.time() - t0:.0f}s", flush=True)
    out["__names__"] = np.asarray(json.dumps(names))
    np.savez_compressed(gg.OUT / "ks2.npz", **out)


if __name__ == "__main__":
    main()
