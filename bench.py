#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched JobShopLab env-step path on B200 (BASELINE.json metric).

A bench "step" is one pass of the hot path over one batch: every env of a synthetic ta41-shape
(30 jobs x 20 machines, 30 AGVs) batch is reset and rolled to the end of its episode by the in-kernel
uniform-random policy (Philox), in ONE launch of k_rollout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU).  Instances shard by env index, no data-path
collective; one small all-reduce of (steps, time) at the end.  `--impl reference` times the CPU oracle
port of the reference (oracle/jsl_oracle.c) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

J, M = 30, 20                      # ta41 shape
BYTES_PER_STEP = 3307              # SURVEY 8(d) / BASELINE.md 4: 2*S + 11 + O for (J,M,T) = (30,20,30)
METRIC = "env_steps_per_sec"
SEED = 0


# ------------------------------------------------------------------------------------------- workload
BLOCK = 4096  # instances are generated in aligned blocks so that instance i never depends on the slice asked for


def make_instances(n: int, first: int = 0):
    """Taillard-style random 30x20 JSSPs: per job a uniform random machine permutation, durations
    uniform in 1..99.  Block k (instances k*BLOCK ..) comes from its own counter-based Philox stream
    (key 1234, counter word 3 = k), so any shard regenerates exactly its own slice."""
    om = np.empty((n, J * M), dtype=np.int32)
    od = np.empty((n, J * M), dtype=np.int32)
    k0, k1 = first // BLOCK, (first + n - 1) // BLOCK
    for k in range(k0, k1 + 1):
        rng = np.random.Generator(np.random.Philox(key=1234, counter=[0, 0, 0, k]))
        keys = rng.random((BLOCK, J, M))
        bm = np.argsort(keys, axis=2).reshape(BLOCK, J * M).astype(np.int32)
        bd = rng.integers(1, 100, size=(BLOCK, J * M), dtype=np.int32)
        lo, hi = max(first, k * BLOCK), min(first + n, (k + 1) * BLOCK)
        om[lo - first:hi - first] = bm[lo - k * BLOCK:hi - k * BLOCK]
        od[lo - first:hi - first] = bd[lo - k * BLOCK:hi - k * BLOCK]
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


# ------------------------------------------------------------------------------------------- CPU leg
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = max(threads * 64, 256)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_run(min(n, threads * 2), threads)
    steps, secs = 0, 0.0
    for k in range(args.steps):
        r, dt = cpu_run(n, threads, first=k * n)
        steps += r["total_steps"]; secs += dt
    v = steps / secs
    sample = f"{n} ta41-shape episodes per step on {threads} host threads (random policy, Philox seed {SEED})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": "ta41-shape (30 jobs x 20 machines, 30 AGVs) synthetic JSSP batch, uniform-random policy, "
                        "full episodes (reset + rollout to done)",
            "envs_per_gpu": args.envs_per_gpu, "global_envs": args.envs_per_gpu * world, "policy": "random",
            "seed": SEED, "parallelism": f"env-sharded x{world}",
            "l2_policy": "inputs larger than L2 (op tables + state image >> 126 MB)"}


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs-per-gpu", type=int, default=262144)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--step-mode-steps", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from jobshoplab_b200 import engine

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
    lb, mat = taillard_bounds(om, od)
    batch = engine.Batch(tmpl, B, device=local, seed=SEED, variants=(om, od, lb, mat), obs_kind=1)
    batch.set_env_base(first)
    pin = [torch.from_numpy(a).pin_memory() for a in (om, od, lb, mat)]
    host_out = [torch.empty(B, dtype=t.dtype).pin_memory() for t in
                (batch.rout.makespan, batch.rout.n_steps, batch.rout.ret, batch.rout.status)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        return batch.rollout("random", reset_first=True, writeback=False)

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

    # ---- e2e: host buffers in, host results out, through the public API -----------------------------
    # Every pass uploads its own inputs (the 1.26 GB of op tables) from pinned host memory and reads its results
    # back.  The upload of pass k+1 goes to staging buffers on a copy stream while pass k's rollout runs; the
    # staged tables are committed (packed into the live tables) on the compute stream between the rollouts.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    packed = torch.cuda.Event()

    def stage():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(packed)      # the previous commit has consumed the staging area
            batch.stage_variants(*pin)

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
    h2d = sum(t.numel() * t.element_size() for t in pin)
    d2h = sum(t.numel() * t.element_size() for t in host_out)

    # ---- step mode (state round-trips HBM every env.step, observation written): secondary figure ----
    sb = engine.Batch(tmpl, B, device=local, seed=SEED, variants=(om, od, lb, mat), obs_kind=1)
    sb.reset()
    acts = torch.randint(0, 2, (args.step_mode_steps + 8, B), dtype=torch.uint8, device=dev)
    for i in range(8):
        sb.step(acts[i])
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.step_mode_steps):
        sb.step(acts[8 + i])
    s1.record()
    barrier()
    step_ms = s0.elapsed_time(s1)
    # ---- instances generated on the device (InstanceRandomizer) instead of uploaded ---------
    def gen_pass(k):
        sb.randomize(seed=SEED + 1 + k, min_duration=1, max_duration=100)
        r = sb.rollout("random", reset_first=True, writeback=False)
        for h, d in zip(host_out, (r.makespan, r.n_steps, r.ret, r.status)):
            h.copy_(d, non_blocking=True)

    gen_pass(0)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    gen_steps = 0
    for k in range(args.steps):
        gen_pass(1 + k)
        torch.cuda.current_stream().synchronize()
        gen_steps += int(host_out[1].sum().item())
    g1.record()
    barrier()
    gen_ms = g0.elapsed_time(g1)
    step_bytes = 2 * sb.shape.state_bytes + 1 + 8 + 2 + sb.shape.obs_bytes + 4 + 4 + 8 + 1
    shape_state_bytes, shape_obs_bytes = sb.shape.state_bytes, sb.shape.obs_bytes
    sb.close()

    # ---- reduce over ranks: max time, summed units ---------------------------------------------------
    vec = torch.tensor([ms, e2e_ms, step_ms, gen_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([steps_per_pass * args.steps, e2e_steps * args.steps, ok, gen_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms, step_ms, gen_ms = vec.tolist()
    total_steps, total_e2e_steps, ok, total_gen_steps = cnt.tolist()

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        value = total_steps / (ms / 1e3)
        k_avg_ms = float(np.mean(kernel_ms))
        achieved = BYTES_PER_STEP * steps_per_pass / (k_avg_ms / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "r01_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        step_rate = B * world * args.step_mode_steps / (step_ms / 1e3)
        out = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
            "per_gpu": value / world, "episodes_terminated": ok, "env_steps_per_pass": total_steps / args.steps,
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": total_e2e_steps / (e2e_ms / 1e3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "what": "per pass: pinned host op tables -> device (jsl_batch_stage_variants on a copy stream, "
                            "overlapping the previous pass's rollout) + jsl_batch_commit_variants + jsl_rollout + D2H of "
                            "makespan/n_steps/return/status; the first timed pass's upload is inside the timed region"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_rollout<1>", "peak_source": peak_src,
                         "bytes_per_env_step": BYTES_PER_STEP,
                         "note": "algorithmic step-mode bytes (SURVEY 8d) x env-steps per launch / launch time; in "
                                 "rollout mode the state stays in shared memory, so real DRAM traffic is far lower "
                                 "(see traffic) and the kernel is issue/latency bound, not HBM bound"},
            "step_mode": {"value": step_rate, "unit": "env-steps/s", "kernel": "k_step<1>",
                          "bytes_per_env_step": step_bytes, "state_bytes": shape_state_bytes,
                          "obs_bytes": shape_obs_bytes,
                          "achieved_gbs": step_bytes * step_rate / world / 1e9,
                          "frac_of_hbm_peak": step_bytes * step_rate / world / 1e9 / peak,
                          "what": "jsl_step on the same batch with device-resident random actions: state image "
                                  "read+written and the default observation written every env.step"},
        }
        out["device_instances"] = {
            "value": total_gen_steps / (gen_ms / 1e3), "unit": "env-steps/s",
            "what": "jsl_batch_randomize (InstanceRandomizer on the device: a fresh 30x20 instance per env) + "
                    "jsl_rollout + D2H of the results, per pass; no host tables"}
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            n = max(threads * 128, 256)
            r, dt = cpu_run(n, threads, first=0)
            gpu_mk = batch.rout.makespan[:n].cpu().numpy() if first == 0 else None
            out["cpu_baseline"] = {
                "value": r["total_steps"] / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
                "sample": f"first {n} envs of the same batch (same instances, same Philox streams) on {threads} host "
                          f"threads, {dt:.1f} s",
                "makespans_equal_gpu": bool(gpu_mk is not None and np.array_equal(gpu_mk, r["makespan"])),
            }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
