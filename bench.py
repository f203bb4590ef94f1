#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched JobShopLab env-step path on B200 (BASELINE.json metric).

A bench "step" is one pass of the hot path over one batch: every env of a synthetic ta41-shape
(30 jobs x 20 machines, 30 AGVs) batch is reset and rolled to the end of its episode by the in-kernel
uniform-random policy (Philox), in ONE launch of k_rollout (BASELINE.json config 3: 1 M instances).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU).  Envs shard by index, no data-path collective; the one
collective north_star names -- the final makespan / return reduction (parallel.reduce_stats: SUM / MIN / MAX
all-reduces over NCCL incl. a 64-bin makespan histogram) -- closes the run and is printed as `makespan_stats`.

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the UNMODIFIED
Python reference (oracle/_ref, installed by oracle/make_ref.py) on one process per core; where it is not
installed, the C oracle port (oracle/jsl_oracle.c).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

J, M = 30, 20                      # ta41 shape
BYTES_PER_STEP = 3307              # SURVEY 8(d) / BASELINE.md 4: 2*S + 11 + O for (J,M,T) = (30,20,30)
METRIC = "env_steps_per_sec"
SEED = 0
CONFIG3_ENVS = 1 << 20             # BASELINE.json config 3: "1M instances"
WORKLOAD = ("ta41-shape (30 jobs x 20 machines, 30 AGVs) synthetic JSSP batch, uniform-random policy, "
            "full episodes (reset + rollout to done)")


# ------------------------------------------------------------------------------------------- workload
BLOCK = 4096  # instances are generated in aligned blocks so that instance i never depends on the slice asked for


def _make_block(k: int):
    rng = np.random.Generator(np.random.Philox(key=1234, counter=[0, 0, 0, k]))
    keys = rng.random((BLOCK, J, M))
    bm = np.argsort(keys, axis=2).reshape(BLOCK, J * M).astype(np.int32)
    bd = rng.integers(1, 100, size=(BLOCK, J * M), dtype=np.int32)
    return bm, bd


def make_instances(n: int, first: int = 0, threads: int | None = None):
    """Taillard-style random 30x20 JSSPs: per job a uniform random machine permutation, durations
    uniform in 1..99.  Block k (instances k*BLOCK ..) comes from its own counter-based Philox stream
    (key 1234, counter word 3 = k), so any shard regenerates exactly its own slice."""
    om = np.empty((n, J * M), dtype=np.int32)
    od = np.empty((n, J * M), dtype=np.int32)
    k0, k1 = first // BLOCK, (first + n - 1) // BLOCK

    def fill(k):
        bm, bd = _make_block(k)
        lo, hi = max(first, k * BLOCK), min(first + n, (k + 1) * BLOCK)
        om[lo - first:hi - first] = bm[lo - k * BLOCK:hi - k * BLOCK]
        od[lo - first:hi - first] = bd[lo - k * BLOCK:hi - k * BLOCK]

    threads = threads or min(16, max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if k1 - k0 < 2 or threads == 1:
        for k in range(k0, k1 + 1):
            fill(k)
    else:  # numpy releases the GIL in the generator and in argsort
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(fill, range(k0, k1 + 1)))
    return om, od


def taillard_bounds(om: np.ndarray, od: np.ndarray):
    """Vectorised jobshoplab/utils/utils.py:310-352 (lower bound, max_allowed_time) for [n][J*M] tables."""
    n = om.shape[0]
    mach = om.reshape(n, J, M)
    dur = od.reshape(n, J, M).astype(np.int64)
    pos = np.argsort(mach, axis=2)                      # pos[i,j,m] = op index of machine m in job j
    csum = np.cumsum(dur, axis=2)
    d_at = np.take_along_axis(dur, pos, axis=2)
    before = np.take_along_axis(csum, pos, axis=2) - d_at
    total = csum[:, :, -1:]
    after = total - before - d_at
    b = before.min(axis=1)
    a = after.min(axis=1)
    T = d_at.sum(axis=1)
    lb = np.maximum((b + T + a).max(axis=1), total[:, :, 0].max(axis=1))
    mat = dur.sum(axis=(1, 2))
    return lb.astype(np.int32), mat.astype(np.int32)


def taillard_bounds_torch(om, od, chunk: int = 1 << 16):
    """The same arithmetic on torch tensors (any device), chunked: host numpy takes a minute for 1 M instances."""
    import torch

    lbs, mats = [], []
    for s in range(0, om.shape[0], chunk):
        mach = om[s:s + chunk].reshape(-1, J, M).long()
        dur = od[s:s + chunk].reshape(-1, J, M).long()
        pos = torch.argsort(mach, dim=2, stable=True)
        csum = torch.cumsum(dur, dim=2)
        d_at = torch.gather(dur, 2, pos)
        before = torch.gather(csum, 2, pos) - d_at
        total = csum[:, :, -1:]
        after = total - before - d_at
        b, a, T = before.min(dim=1).values, after.min(dim=1).values, d_at.sum(dim=1)
        lbs.append(torch.maximum((b + T + a).max(dim=1).values, total[:, :, 0].max(dim=1).values).int())
        mats.append(dur.sum(dim=(1, 2)).int())
    return torch.cat(lbs), torch.cat(mats)


def template_instance():
    """A 30x20 classic instance (default buffers/AGVs of the reference) used as the shape template."""
    from jobshoplab_b200 import compiler

    om, od = make_instances(1)
    lines = [f"{J} {M}"]
    for j in range(J):
        lines.append(" ".join(f"{om[0, j * M + k]} {od[0, j * M + k]}" for k in range(M)))
    return compiler.compile_spec_text("\n".join(lines) + "\n")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU legs
def cpu_run(n_envs: int, threads: int, first: int = 0):
    """Oracle port on host threads over the same instances / Philox streams as the GPU arm."""
    from oracle import oracle_py as op
    from jobshoplab_b200 import abi
    import copy

    tmpl = template_instance()
    om, od = make_instances(n_envs, first)
    lb, mat = taillard_bounds(om, od)
    insts = []
    for i in range(n_envs):
        li = copy.copy(tmpl)
        li.op_machine, li.op_duration = om[i].copy(), od[i].copy()
        li.lower_bound, li.max_allowed_time = int(lb[i]), int(mat[i])
        insts.append(li)
    op.lib()
    t0 = time.perf_counter()
    r = op.rollout_batch(insts, n_envs, abi.POLICY_RANDOM, seed=SEED, n_threads=threads, env_base=first)
    dt = time.perf_counter() - t0
    return r, dt


def reference_pool(workers: int, episodes_per_worker: int = 4, first: int = 0):
    """`workers` processes, each owning one UNMODIFIED reference env (oracle/ref_bench.py) over the bench's own
    instances first + w, first + w + workers, ... and the GPU arm's own Philox action streams."""
    from oracle import ref_bench as rb

    n = workers * episodes_per_worker
    om, od = make_instances(n, first)
    work = [[(first + w + k * workers, rb.spec_text(om[w + k * workers], od[w + k * workers], J, M))
             for k in range(episodes_per_worker)] for w in range(workers)]
    return rb.ReferencePool(work, seed=SEED)


def reference_baseline(serial_steps: int = 1500, pool_steps: int = 0):
    """The Python reference, serially and on one process per host core (SURVEY 8(d) CPU baseline): only env.step
    is timed.  Returns the cpu_baseline fields and the finished episodes [(env index, makespan, n_steps)]."""
    from oracle import ref_bench as rb

    cores = os.cpu_count() or 1
    p1 = reference_pool(1)
    p1.run(50)                                       # warm-up
    s = p1.run(serial_steps)
    s_tail = p1.run(0)                               # finish env 0's episode (timed too: same loop)
    p1.close()
    pn = reference_pool(cores)
    pn.run(20)
    p = pn.run(pool_steps)                           # <= 0: every worker finishes its own episode
    pn.close()
    episodes = s["episodes"] + s_tail["episodes"] + p["episodes"]
    serial = (s["steps"] + s_tail["steps"]) / (s["seconds"] + s_tail["seconds"])
    return {"value": p["steps"] / p["seconds"], "unit": "env-steps/s", "cores": cores, "kind": "reference",
            "serial": serial, "pool": p["steps"] / p["seconds"], "cpu_model": rb.cpu_model(),
            "instance": "the bench's own ta41-shape instances (env i = instance i)", "policy": "random (the GPU arm's Philox streams)",
            "sample": f"unmodified Python reference (JobShopLabEnv, Config() defaults, SpecRepository), env.step loop only: "
                      f"serial {s['steps'] + s_tail['steps']} steps of env 0 in {s['seconds'] + s_tail['seconds']:.1f} s; "
                      f"pool {cores} processes, one whole episode each (envs 0..{cores - 1}, {p['steps']} steps) in {p['seconds']:.1f} s"}, episodes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import ref_bench as rb

    if rb.available():
        # each bench step: every host core advances its own reference env by `per` env.step calls
        per = 250
        pool = reference_pool(threads, episodes_per_worker=8)
        for _ in range(args.warmup):
            pool.run(per)
        steps, secs = 0, 0.0
        for _ in range(args.steps):
            r = pool.run(per)
            steps += r["steps"]; secs += r["seconds"]
        pool.close()
        v = steps / secs
        kind = "reference"
        sample = (f"unmodified Python reference (oracle/_ref): {threads} processes (one per host thread), each stepping its "
                  f"own ta41-shape env {per} env.step calls per bench step with the GPU arm's Philox actions; only "
                  f"env.step is timed, rate = sum of steps / slowest worker's in-loop seconds")
        cfg = {"workload": WORKLOAD, "envs": threads, "env_steps_per_bench_step": per * threads, "policy": "random",
               "seed": SEED, "parallelism": f"{threads} host processes", "cpu_model": rb.cpu_model()}
    else:
        n = max(threads * 64, 256)
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_run(min(n, threads * 2), threads)
        steps, secs = 0, 0.0
        for k in range(args.steps):
            r, dt = cpu_run(n, threads, first=k * n)
            steps += r["total_steps"]; secs += dt
        v = steps / secs
        kind = "port"
        sample = (f"C oracle port (oracle/_ref absent): {n} ta41-shape episodes per step on {threads} host threads "
                  f"(random policy, Philox seed {SEED})")
        cfg = {"workload": WORKLOAD, "envs": n, "policy": "random", "seed": SEED, "parallelism": f"{threads} host threads"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": WORKLOAD,
            "envs_per_gpu": args.envs_per_gpu, "global_envs": args.envs_per_gpu * world, "policy": "random",
            "seed": SEED, "parallelism": f"env-sharded x{world}",
            "l2_policy": "inputs larger than L2 (op tables + state image >> 126 MB)"}


# ------------------------------------------------------------------------------------------- secondary configs
def timed_rollouts(torch, batch, policy, reps=3, warm=1):
    for _ in range(warm):
        batch.rollout(policy, reset_first=True, writeback=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = batch.rollout(policy, reset_first=True, writeback=False)
    e1.record()
    torch.cuda.synchronize()
    return r, int(r.n_steps.sum()) * reps, e0.elapsed_time(e1) / 1e3


def secondary_configs(torch, dist, dev, local, rank, world, B2: int):
    """BASELINE.json configs 1, 2, 4, 5 in compact form (the headline is config 3).  Configs 1/2/4 run on rank 0;
    config 5 is sharded over all ranks and its makespan distribution is all-reduced and KS-tested."""
    sys.path.insert(0, str(ROOT / "tests"))
    import _golden as G
    from jobshoplab_b200 import compiler, engine, parallel
    from jobshoplab_b200.env import JobShopLabEnv

    out = []
    z = np.load(ROOT / "tests" / "golden" / "rules_all.npz", allow_pickle=False)
    names, texts = json.loads(str(z["names"])), json.loads(str(z["texts"]))
    table = z["table"]
    if rank == 0:
        # config 1: single JobShopLabEnv on ft06, random actions (the reference's getting-started loop)
        env = JobShopLabEnv(compiler=compiler.compile_spec_text(texts[names.index("spec/ft06")]), seed=1)
        rng = np.random.default_rng(0)
        n, t0 = 0, time.perf_counter()
        for _ in range(5):
            env.reset()
            while not env.done:
                env.step(int(rng.integers(0, 2)))
                n += 1
        dt = time.perf_counter() - t0
        env.close()
        out.append({"config": 1, "workload": "ft06 single JobShopLabEnv (drop-in API, host loop), random actions, 5 episodes",
                    "value": n / dt, "unit": "env-steps/s"})
        # config 2: ft10 / la16 x B2, SPT / MWKR rollouts, bit-exact makespans
        for inst in ("ft10", "la16"):
            i = names.index(f"spec/{inst}")
            batch = engine.Batch(compiler.compile_spec_text(texts[i]), B2, device=local)
            for r_i, pol in ((1, "spt"), (2, "mwkr")):
                r, steps, secs = timed_rollouts(torch, batch, pol)
                out.append({"config": 2, "workload": f"{inst} x{B2} {pol} rollout", "value": steps / secs, "unit": "env-steps/s",
                            "makespan": int(r.makespan[0]), "bit_exact_vs_reference": bool((r.makespan == int(table[i, r_i, 0])).all())})
            batch.close()
        # config 4: transport instances with FIFO pre/post buffers of capacity 5, random actions
        for name in ("la01-t/fifo5/random0", "ft06-t/fifo5/random0"):
            sc = G.Scenario("buffers", name)
            batch = engine.Batch(sc.lowered, B2, device=local, seed=sc.meta["seed"], **sc.cfg)
            r, steps, secs = timed_rollouts(torch, batch, "random")
            st = r.status
            e0 = sc.meta["env_idx"]
            out.append({"config": 4, "workload": f"{name.rsplit('/', 1)[0]} (FIFO buffers, capacity 5) x{B2} random rollout",
                        "value": steps / secs, "unit": "env-steps/s", "terminated": int((st == 1).sum()),
                        "buffer_full": int((st == 4).sum()),
                        "bit_exact_vs_reference": bool(int(r.makespan[e0]) == sc.meta["makespan"] and int(r.n_steps[e0]) == sc.meta["n_steps"])})
            batch.close()
    # config 5: stochastic ft20-t (setup times, gamma breakdowns, Poisson travel), sharded over the ranks
    fx = G.KsFixture("ft20-t_features")
    batch = engine.Batch(fx.lowered, B2, device=local, seed=5)
    batch.set_env_base(rank * B2)
    batch.rollout("fifo", reset_first=True, writeback=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = batch.rollout("fifo", reset_first=True, writeback=False)
    e1.record()
    torch.cuda.synchronize()
    secs = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    lo, hi, bins = 0, 4096, 4096  # one bin per integer makespan: the reduced histogram IS the empirical distribution
    st = parallel.reduce_stats(r.makespan, r.n_steps, r.ret, r.status, bins=bins, lo=lo, hi=hi)
    if rank == 0:
        from scipy import stats as sps

        hist = st.histogram.numpy()
        ref = np.sort(fx.makespan)
        # two-sample KS statistic between the all-reduced histogram and the reference sample
        edges = np.arange(lo, hi + 1)
        cdf_gpu = np.cumsum(hist) / max(hist.sum(), 1)
        cdf_ref = np.searchsorted(ref, edges[1:] - 0.5, side="right") / len(ref)
        d = float(np.abs(cdf_gpu - cdf_ref).max())
        n1, n2 = int(hist.sum()), len(ref)
        p = float(sps.kstwo.sf(d, int(round(n1 * n2 / (n1 + n2)))))
        out.append({"config": 5, "workload": f"stochastic ft20-t (setup times, gamma breakdowns, Poisson travel) x{B2} per GPU, "
                                             f"always-accept rollout, sharded x{world}",
                    "value": st.steps * 3 / float(secs.item()), "unit": "env-steps/s", "episodes": st.count,
                    "terminated": st.terminated, "makespan_mean": st.makespan_mean, "reference_makespan_mean": float(ref.mean()),
                    "ks_statistic": d, "ks_pvalue": p, "ks_alpha": 0.01, "ks_pass": bool(p > 0.01),
                    "reference_sample": n2, "reduced_over": f"NCCL all-reduce x{world}" if world > 1 else "single rank"})
    batch.close()
    return out


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs-per-gpu", type=int, default=CONFIG3_ENVS)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--step-mode-steps", type=int, default=1000)
    ap.add_argument("--secondary-envs", type=int, default=65536)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from jobshoplab_b200 import engine, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.envs_per_gpu
    first = rank * B

    tmpl = template_instance()
    om, od = make_instances(B, first)
    pin = [torch.from_numpy(om).pin_memory(), torch.from_numpy(od).pin_memory()]
    del om, od                                   # (the pinned copies are the host-side inputs from here on)
    om_d, od_d = pin[0].to(dev, non_blocking=True), pin[1].to(dev, non_blocking=True)
    lb_d, mat_d = taillard_bounds_torch(om_d, od_d)
    pin += [lb_d.cpu().pin_memory(), mat_d.cpu().pin_memory()]
    del om_d, od_d, lb_d, mat_d
    batch = engine.Batch(tmpl, B, device=local, seed=SEED, variants=tuple(t.numpy() for t in pin), obs_kind=1)
    batch.set_env_base(first)
    host_out = [torch.empty(B, dtype=t.dtype).pin_memory() for t in
                (batch.rout.makespan, batch.rout.n_steps, batch.rout.ret, batch.rout.status)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(b=batch):
        return b.rollout("random", reset_first=True, writeback=False)

    for _ in range(args.warmup):
        one_pass()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = batch.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        one_pass()
        ev[k + 1].record()
    barrier()
    clk = clocks.stop()
    launches = batch.launches - launches0
    ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    steps_per_pass = int(batch.rout.n_steps.sum().item())
    ok = int((batch.rout.status == 1).sum().item())
    # the collective north_star names: final makespan / return reduction over NCCL (tiny; outside the timed region)
    r = batch.rout
    stats = parallel.reduce_stats(r.makespan, r.n_steps, r.ret, r.status, bins=64, lo=0, hi=8192)
    gpu_first_mk = batch.rout.makespan[:256].cpu().numpy() if first == 0 else None
    gpu_first_ns = batch.rout.n_steps[:256].cpu().numpy() if first == 0 else None

    # ---- e2e: host buffers in, host results out, through the public API -----------------------------
    # Every pass uploads its own inputs (the op tables) from pinned host memory and reads its results
    # back.  The upload of pass k+1 goes to staging buffers on a copy stream while pass k's rollout runs; the
    # staged tables are committed (packed into the live tables) on the compute stream between the rollouts.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    packed = torch.cuda.Event()

    # host-side input format of the e2e path: uint8 machines + uint16 durations (3 B per operation instead of 8)
    pin_e2e = list(engine.Batch.pack_variants(pin[0], pin[1])) + pin[2:]

    def stage():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(packed)      # the previous commit has consumed the staging area
            batch.stage_variants_packed(*pin_e2e)

    def e2e_pass(more):
        main_stream.wait_stream(copy_stream)    # this pass's inputs are on the device
        batch.commit_variants()
        packed.record(main_stream)
        if more:
            stage()                             # next pass's upload overlaps this pass's rollout
        r = batch.rollout("random", reset_first=True, writeback=False)
        for h, d in zip(host_out, (r.makespan, r.n_steps, r.ret, r.status)):
            h.copy_(d, non_blocking=True)

    packed.record(main_stream)
    stage()
    e2e_pass(False)                             # untimed warm-up pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    copy_stream.wait_event(e0)
    stage()                                     # the first timed pass's upload: nothing to hide it behind
    for k in range(args.steps):
        e2e_pass(k + 1 < args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_steps = int(host_out[1].sum().item())
    assert e2e_steps == steps_per_pass, "the pipelined upload changed the instances"  # same tables, same streams
    h2d = sum(t.numel() * t.element_size() for t in pin_e2e)
    d2h = sum(t.numel() * t.element_size() for t in host_out)

    # ---- strong scaling: BASELINE config 3's 1 M envs split over the ranks ------------------------------------
    strong = None
    if world > 1:
        gs = CONFIG3_ENVS
        sf, sc = parallel.shard_range(gs, rank, world)
        s_om, s_od = make_instances(sc, sf)
        s_lb, s_mat = taillard_bounds_torch(torch.from_numpy(s_om).to(dev), torch.from_numpy(s_od).to(dev))
        sbatch = engine.Batch(tmpl, sc, device=local, seed=SEED, variants=(s_om, s_od, s_lb.cpu().numpy(), s_mat.cpu().numpy()), obs_kind=1)
        sbatch.set_env_base(sf)
        for _ in range(2):
            one_pass(sbatch)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            one_pass(sbatch)
        q1.record()
        barrier()
        strong = (q0.elapsed_time(q1), int(sbatch.rout.n_steps.sum().item()) * args.steps)
        sbatch.close()

    # ---- step mode (state round-trips HBM every env.step, observation written): the kernel that moves its bytes ----
    batch.reset()
    pool_rows = 128
    acts = torch.randint(0, 2, (pool_rows, B), dtype=torch.uint8, device=dev)
    for i in range(8):
        batch.step(acts[i])
    barrier()
    sl0 = batch.launches
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.step_mode_steps):
        batch.step(acts[(8 + i) % pool_rows])
    s1.record()
    barrier()
    step_ms = s0.elapsed_time(s1)
    step_launches = batch.launches - sl0
    step_live = int((batch.out.status == 0).sum().item())  # envs still mid-episode after the timed steps
    # ---- instances generated on the device (InstanceRandomizer) instead of uploaded ---------
    def gen_pass(k):
        batch.randomize(seed=SEED + 1 + k, min_duration=1, max_duration=100)
        r = batch.rollout("random", reset_first=True, writeback=False)
        for h, d in zip(host_out, (r.makespan, r.n_steps, r.ret, r.status)):
            h.copy_(d, non_blocking=True)

    gen_reps = min(args.steps, 3)
    gen_pass(0)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    gen_steps = 0
    for k in range(gen_reps):
        gen_pass(1 + k)
        torch.cuda.current_stream().synchronize()
        gen_steps += int(host_out[1].sum().item())
    g1.record()
    barrier()
    gen_ms = g0.elapsed_time(g1)
    shape_state_bytes, shape_obs_bytes = batch.shape.state_bytes, batch.shape.obs_bytes
    step_bytes = 2 * shape_state_bytes + 1 + 8 + 2 + shape_obs_bytes + 4 + 4 + 8 + 1
    batch.close()
    del batch

    secondary = None
    if not args.no_secondary:
        secondary = secondary_configs(torch, dist, dev, local, rank, world, args.secondary_envs)

    # ---- reduce over ranks: max time, summed units ---------------------------------------------------
    vec = torch.tensor([ms, e2e_ms, step_ms, gen_ms, strong[0] if strong else 0.0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([steps_per_pass * args.steps, e2e_steps * args.steps, ok, gen_steps, strong[1] if strong else 0,
                        step_live], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms, step_ms, gen_ms, strong_ms = vec.tolist()
    total_steps, total_e2e_steps, ok, total_gen_steps, strong_steps, step_live = cnt.tolist()

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        value = total_steps / (ms / 1e3)
        k_avg_ms = float(np.mean(kernel_ms))
        rollout_equiv = BYTES_PER_STEP * steps_per_pass / (k_avg_ms / 1e3) / 1e9
        stored = {}
        tp = ROOT / "profiles" / "r02_traffic.json"
        if tp.exists():
            stored = json.loads(tp.read_text())
        step_rate = B * world * args.step_mode_steps / (step_ms / 1e3)
        step_launch_ms = step_ms / max(step_launches, 1)
        achieved = BYTES_PER_STEP * B / (step_launch_ms / 1e3) / 1e9
        moved = step_bytes * B / (step_launch_ms / 1e3) / 1e9
        sm_mhz = clk.get("sm_mhz") or 1965.0
        ipe = float(stored.get("rollout_warp_instr_per_env_step", 510.0))
        ceiling = 148 * 4 * sm_mhz * 1e6 / ipe
        out = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
            "per_gpu": value / world, "episodes_terminated": ok, "env_steps_per_pass": total_steps / args.steps,
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": total_e2e_steps / (e2e_ms / 1e3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "what": "per pass: pinned host op tables (uint8 machines, uint16 durations) -> device (jsl_batch_stage_variants_packed on a copy stream, "
                            "overlapping the previous pass's rollout) + jsl_batch_commit_variants + jsl_rollout + D2H of "
                            "makespan/n_steps/return/status; the first timed pass's upload is inside the timed region"},
            "roofline": {
                "bound": "hbm", "kernel": "k_step<1,0,2>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "bytes_per_env_step": BYTES_PER_STEP,
                "launch_ms": step_launch_ms, "env_steps_per_launch": B,
                "traffic": stored.get("step_dram_bytes_per_env_step", None) and stored["step_dram_bytes_per_env_step"] * B,
                "traffic_source": "stored: profiles/r02_traffic.json (ncu dram__bytes_read+write of the same kernel and "
                                  "shape, per env-step x envs per launch), not measured in this run",
                "moved": {"bytes_per_env_step": step_bytes, "state_image_bytes": shape_state_bytes, "obs_bytes": shape_obs_bytes,
                          "gbs": moved, "frac": moved / peak},
                "note": "the step-mode kernel is the one that really moves its bytes (state image in + out and the "
                        "observation every env.step): `achieved` = SURVEY 8(d)'s 3307 algorithmic B x envs per launch / "
                        "the launch time measured here with CUDA events over the timed jsl_step launches; `moved` "
                        "uses the bytes of the implemented layout",
                "rollout_equivalent": {
                    "kernel": "k_rollout<1,0,2>", "achieved": rollout_equiv, "frac": rollout_equiv / peak,
                    "note": "normalised throughput of the headline kernel, NOT a bandwidth: 3307 B x env-steps per launch / "
                            "launch time; in rollout mode the state never leaves shared memory (stored ncu traffic: "
                            f"{stored.get('rollout_dram_bytes_per_env_step', 4.1)} B per env-step)"}},
            "issue_roofline": {
                "kernel": "k_rollout<1,0,2>", "bound": "instruction issue", "instr_per_env_step": ipe,
                "instr_source": "stored: profiles/r02_traffic.json (ncu smsp__inst_executed / env-steps)",
                "sm_mhz": sm_mhz, "ceiling": ceiling, "achieved": value / world, "frac": value / world / ceiling,
                "unit": "env-steps/s/GPU", "note": "148 SMs x 4 schedulers x SM clock / warp-instructions per env-step"},
            "step_mode": {"value": step_rate, "unit": "env-steps/s", "kernel": "k_step<1,0,2>",
                          "bytes_per_env_step": step_bytes, "state_bytes": shape_state_bytes,
                          "obs_bytes": shape_obs_bytes, "achieved_gbs": moved, "frac_of_hbm_peak": moved / peak,
                          "algorithmic_frac_of_hbm_peak": BYTES_PER_STEP * step_rate / world / 1e9 / peak,
                          "envs_still_running": step_live,
                          "what": "jsl_step on the same batch with device-resident random actions: state image "
                                  "read+written and the default observation written every env.step"},
            "makespan_stats": {"count": stats.count, "terminated": stats.terminated, "errors": stats.errors,
                               "mean": stats.makespan_mean, "std": stats.makespan_std, "min": stats.makespan_min,
                               "max": stats.makespan_max, "return_mean": stats.return_sum / max(stats.count, 1),
                               "histogram_64_bins_0_8192": stats.histogram.tolist(),
                               "how": f"parallel.reduce_stats over {'NCCL x%d' % world if world > 1 else 'one rank'}"},
            "strong": {"global_envs": CONFIG3_ENVS,
                       "value": (strong_steps / (strong_ms / 1e3)) if world > 1 else value,
                       "ms_per_step": (strong_ms / args.steps) if world > 1 else ms / args.steps,
                       "what": "BASELINE config 3 as stated: the same 1 M envs split over the ranks (at N=1 this IS the headline run)"},
        }
        out["device_instances"] = {
            "value": total_gen_steps / (gen_ms / 1e3), "unit": "env-steps/s",
            "what": "jsl_batch_randomize (InstanceRandomizer on the device: a fresh 30x20 instance per env) + "
                    "jsl_rollout + D2H of the results, per pass; no host tables"}
        if secondary is not None:
            out["secondary"] = secondary
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            n = max(threads * 128, 256)
            r, dt = cpu_run(n, threads, first=0)
            port = {"value": r["total_steps"] / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
                    "sample": f"C oracle port: first {n} envs of the same batch (same instances, same Philox streams) on "
                              f"{threads} host threads, {dt:.1f} s",
                    "makespans_equal_gpu": bool(gpu_first_mk is not None and np.array_equal(gpu_first_mk, r["makespan"][:256]))}
            from oracle import ref_bench as rb
            if rb.available():
                cb, episodes = reference_baseline()
                cb["port"] = port
                cb["episodes_checked_vs_gpu"] = len(episodes)
                cb["makespans_equal_gpu"] = bool(len(episodes) > 0 and all(
                    i < 256 and term and int(gpu_first_mk[i]) == mk and int(gpu_first_ns[i]) == ns for i, mk, ns, term in episodes))
                out["cpu_baseline"] = cb
            else:
                port["note"] = "oracle/_ref (the Python reference) is not installed here: port only"
                out["cpu_baseline"] = port
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
