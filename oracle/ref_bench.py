"""CPU baseline proper: the UNMODIFIED reference env (oracle/_ref, else /root/reference) stepped on host cores.

BASELINE INFRASTRUCTURE ONLY -- used by bench.py's cpu_baseline leg and `bench.py --impl reference`, never by the
product.  SURVEY 8(d) / BASELINE.md 3: one process = one `JobShopLabEnv` built through the reference's own public
API (`Config()` defaults, `Compiler` + `SpecRepository`), only the `env.step(a)` calls are timed (construction and
`reset` are outside the timer), serial = 1 worker, pool = one worker per host core, all stepping at the same time;
rate = sum of env steps / the slowest worker's in-loop seconds.

The instances are the bench's own synthetic ta41-shape instances (written as Taillard spec files) and the actions
the GPU arm's own streams -- `Philox4x32-10(key=seed, counter=(step, 0, env_lo, env_hi))[0] & 1` -- so every
finished reference episode has a makespan the GPU arm must reproduce exactly (bench.py checks it in the same run).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import tempfile
import time
from pathlib import Path

_HERE = Path(__file__).resolve().parent


def ref_root() -> Path | None:
    env = os.environ.get("JSL_REFERENCE_ROOT")
    for p in ([Path(env)] if env else []) + [_HERE / "_ref", Path("/root/reference")]:
        if (p / "jobshoplab" / "__init__.py").exists():
            return p
    return None


def available() -> bool:
    return ref_root() is not None


def _activate():
    root = ref_root()
    if root is None:
        raise RuntimeError("the reference is not installed: run `python oracle/make_ref.py` where /root/reference exists")
    shims = str(_HERE / "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    if str(root) not in sys.path:
        sys.path.insert(1, str(root))
    return root


def spec_text(om_row, od_row, J: int, M: int) -> str:
    """Taillard spec-file text (`J M` header, then `machine duration ...` per job) of one bench instance."""
    lines = [f"{J} {M}"]
    for j in range(J):
        lines.append(" ".join(f"{int(om_row[j * M + k])} {int(od_row[j * M + k])}" for k in range(M)))
    return "\n".join(lines) + "\n"


def philox_bit(seed: int, env_idx: int, step: int) -> int:
    """Bit 0 of Philox4x32-10(counter=(step, 0, env_lo, env_hi), key=(seed_lo, seed_hi))[0] (Salmon et al. 2011):
    the RANDOM policy of the engine and of the C oracle."""
    c = [step & 0xFFFFFFFF, 0, env_idx & 0xFFFFFFFF, (env_idx >> 32) & 0xFFFFFFFF]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c[0] & 1


def _worker(conn, work, seed):
    """work: list of (env_idx, spec text); episodes are run in that order, cyclically."""
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # the reference's loggers write to stdout
    _activate()
    from jobshoplab import JobShopLabEnv
    from jobshoplab.compiler import Compiler
    from jobshoplab.compiler.repos import SpecRepository
    from jobshoplab.types.config_types import Config

    tmp = tempfile.TemporaryDirectory()
    cur = {"k": -1, "env": None, "idx": None, "step": 0}

    def next_env():
        cur["k"] += 1
        idx, text = work[cur["k"] % len(work)]
        path = Path(tmp.name) / f"inst_{idx}"
        path.write_text(text)
        cfg = Config()
        repo = SpecRepository(path, loglevel="warning", config=cfg)
        cur["env"] = JobShopLabEnv(config=cfg, compiler=Compiler(cfg, loglevel="warning", repo=repo))
        cur["idx"], cur["step"] = idx, 0

    next_env()
    conn.send(("ready", None))
    while True:
        cmd, n = conn.recv()
        if cmd == "stop":
            break
        # cmd == "run": n env steps (n <= 0: to the end of the current episode), finished episodes reported
        steps, secs, episodes = 0, 0.0, []
        whole_episode = n <= 0
        while whole_episode or steps < n:
            env = cur["env"]
            budget = (1 << 60) if whole_episode else n - steps
            idx, k = cur["idx"], cur["step"]
            t0 = time.perf_counter()
            done = 0
            while not env.done and done < budget:
                env.step(philox_bit(seed, idx, k))
                k += 1
                done += 1
            secs += time.perf_counter() - t0
            steps += done
            cur["step"] = k
            if env.done:
                episodes.append((idx, int(env.state.state.time.time), k, bool(env.terminated)))
                next_env()  # untimed: construction + reset
                if whole_episode:
                    break
        conn.send(("done", (steps, secs, episodes)))
    conn.close()


class ReferencePool:
    """`n_workers` processes, each owning one reference env."""

    def __init__(self, work_per_worker, seed: int = 0):
        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for work in work_per_worker:
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, work, seed), daemon=True)
            p.start()
            self.procs.append(p); self.conns.append(a)
        for c in self.conns:
            tag, _ = c.recv()
            assert tag == "ready"

    def run(self, n_steps: int) -> dict:
        """Every worker does n_steps env.step calls (<= 0: finishes its current episode), all at the same time."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("run", n_steps))
        res = [c.recv()[1] for c in self.conns]
        wall = time.perf_counter() - t0
        return {"steps": sum(r[0] for r in res), "seconds": max(r[1] for r in res), "wall": wall,
                "episodes": [e for r in res for e in r[2]]}

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop", 0))
            except OSError:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"
