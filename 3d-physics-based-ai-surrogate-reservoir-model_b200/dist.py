"""Multi-GPU plumbing for the physics loss: the sample axis shards, nothing else does.

Every sample b = (realisation, time point) is an independent unit: the stencil
couples cells of one sample only, dt1/dt2/mbc are per-sample, and the loss terms are sums over
samples.  So ranks take disjoint sample ranges (whole realisations, so kx is not duplicated), run
the kernels with no data-path exchange, and all-reduce only the 16-float loss-term vector.  The
backward needs no collective: the upstream weights are global scalars and each rank's cotangents
are local.  (All-reducing the *network* gradients belongs to the training framework.)
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_realisations(K: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) realisation range of `rank`; contiguous, sizes differ by at most one, no realisation split."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_samples(K: int, T: int, rank: int, world: int) -> Tuple[int, int]:
    """sample range for realisation-major ordering b = r*T + t"""
    lo, hi = shard_realisations(K, rank, world)
    return lo * T, hi * T


def allreduce_terms(terms: torch.Tensor, group=None) -> torch.Tensor:
    """sum the [2][8] (SSE, count) vector over ranks; in place, returns it.  128 bytes: latency bound."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
    return terms


class _Done:
    def wait(self):
        return True


def allreduce_terms_async(terms: torch.Tensor, group=None):
    """Same reduction, not ordered before the work queued next on the current stream: the backward does not read the
    reduced terms (its upstream weights are given), so its kernels need not wait for the 128-byte collective.  Returns
    a handle; ``handle.wait()`` orders the current stream after the reduction (no host synchronisation)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return _Done()
