"""Multi-GPU layout of the env-step path: envs are independent, so they shard by contiguous index range
with NO data-path collective while stepping; one small all-reduce of summary statistics (count, sum,
sum of squares, min, max of the makespans, returns, outcome counts and a fixed-bin makespan histogram)
closes a rollout.  One process per GPU over torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, first+count) of `total` envs owned by `rank`; remainders go to the low ranks."""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


@dataclass
class RolloutStats:
    count: int
    terminated: int
    truncated: int
    errors: int
    steps: int
    makespan_sum: float
    makespan_sqsum: float
    makespan_min: int
    makespan_max: int
    return_sum: float
    histogram: torch.Tensor  # int64 [bins]

    @property
    def makespan_mean(self):
        return self.makespan_sum / max(self.terminated, 1)

    @property
    def makespan_std(self):
        n = max(self.terminated, 1)
        return max(self.makespan_sqsum / n - (self.makespan_sum / n) ** 2, 0.0) ** 0.5


def local_stats(makespan, n_steps, ret, status, bins: int = 64, lo: int = 0, hi: int = 8192):
    """Summary vectors of one rank's rollout results (device tensors) -> (sums f64[6+bins], mins i64[1], maxs i64[1])."""
    ok = status == 1
    mk = makespan[ok].to(torch.float64)
    hist = torch.histc(mk, bins=bins, min=lo, max=hi) if mk.numel() else torch.zeros(bins, dtype=torch.float64, device=makespan.device)
    sums = torch.stack([
        torch.tensor(float(status.numel()), dtype=torch.float64, device=makespan.device),
        ok.sum().to(torch.float64), ((status == 2) | (status == 3)).sum().to(torch.float64), (status > 3).sum().to(torch.float64),
        n_steps.sum().to(torch.float64), mk.sum(), (mk * mk).sum(), ret.sum().to(torch.float64)])
    sums = torch.cat([sums, hist.to(torch.float64)])
    big = torch.iinfo(torch.int64).max
    mn = mk.min().to(torch.int64).reshape(1) if mk.numel() else torch.full((1,), big, dtype=torch.int64, device=makespan.device)
    mx = mk.max().to(torch.int64).reshape(1) if mk.numel() else torch.full((1,), -1, dtype=torch.int64, device=makespan.device)
    return sums, mn, mx


def reduce_stats(makespan, n_steps, ret, status, group=None, bins: int = 64, lo: int = 0, hi: int = 8192) -> RolloutStats:
    """The one collective of a sharded rollout: SUM / MIN / MAX all-reduces of a few hundred bytes."""
    sums, mn, mx = local_stats(makespan, n_steps, ret, status, bins, lo, hi)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    s = sums.tolist()
    return RolloutStats(count=int(s[0]), terminated=int(s[1]), truncated=int(s[2]), errors=int(s[3]), steps=int(s[4]),
                        makespan_sum=s[5], makespan_sqsum=s[6], return_sum=s[7], makespan_min=int(mn.item()),
                        makespan_max=int(mx.item()), histogram=sums[8:].to(torch.int64).cpu())
