"""Multi-GPU plumbing for the pooling path.  Samples are independent (ranks_bev carries the
batch offset, mmdet3d/models/necks/view_transformer.py:246), so the path shards by sample with
NO collective on the data path; torch.distributed is used only to gather per-rank results or
checksums for verification and timings for reporting."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_samples, rank, world):
    """Contiguous, balanced [lo, hi) of the samples rank `rank` pools (first n % world ranks get
    one more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_samples), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_samples(local, n_samples, group=None):
    """All ranks' pooled samples concatenated in sample order (verification only).  `local` is
    this rank's (hi - lo, ...) tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_samples, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def max_over_ranks(values, device=None, group=None):
    """Element-wise max of a list of floats over all ranks (device timings are reported as the
    slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t]
