"""Data parallelism for the hot path: environments shard across ranks (one process per GPU,
torch.distributed over NCCL/NVLink); the only exchanges per update are

  1. all-reduce(sum) of the flat fp32 gradient (88 401 floats at h=196, L=2), every rank having
     normalised its loss by the GLOBAL sample count, so the sum is the global-mean gradient the
     reference's `.mean()` (train.py:554) would give on the concatenated batch;
  2. all-reduce(sum) of 3 float64 scalars {sum G, sum G^2, N} so the return-to-go EMA moments
     (train.py:738-739, 898-901) -- and therefore the next step's normalisation -- match the
     single-GPU result.

Rollout, shaping, scan and per-sample loss need no communication (episodes never span ranks;
Philox counters are keyed by the global env id, so results do not depend on the rank count).
The reference has no multi-GPU path; this is new design, kept host-side and tiny on purpose.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


COLLECTIVES = True      # bench.py switches the all-reduces off to time a rank alone on its shard


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous env-id range [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a small float64 vector over ranks (in place); no-op without a process group."""
    if world()[1] > 1 and COLLECTIVES:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class FlatGradBucket:
    """One flat fp32 buffer holding every gradient: `p.grad` of each parameter IS a view into it, so the kernels and
    autograd accumulate straight into the bucket and an update costs one all-reduce (latency-bound, 353 604 B at
    h=196, L=2), one norm and one scale -- no per-parameter copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def numel(self) -> int:
        return self.flat.numel()

    def attach(self) -> None:
        """(Re)point every p.grad at its slice of the flat buffer (after anything that set grads to None)."""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v

    def zero(self) -> None:
        self.attach()
        self.flat.zero_()

    def allreduce(self, group=None) -> None:
        self.attach()
        if world()[1] > 1 and COLLECTIVES:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def clip_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ (train.py:561) on the flat buffer: returns the total norm before clipping."""
        total = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total
