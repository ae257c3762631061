"""Multi-GPU sampling: shard the batch of shapes across ranks, gather the finished points once.

Shapes are independent (no cross-sample op in the denoiser or the DDPM update), so the path
shards by batch with NO data-path collective; the only exchange is one all-gather of the final
(B/W, N, 3) points (786 KB per rank at B=256, W=8) -- reference analogue: per-rank seeding
`seed + local_rank` (runner/runner.py:39) and the unused all_gather helper (utils/dist_utils.py:58-62).
One process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """Contiguous, balanced [lo, hi) of `total` shapes for `rank` (first `total % W` ranks get one more)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shapes(local, total, group=None):
    """All-gather per-rank (b_r, ...) tensors into the (total, ...) batch in rank order.
    Uses a single all_gather_into_tensor when shards are equal, padded all_gather otherwise."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    W = dist.get_world_size(group)
    sizes = [shard_range(total, r, W) for r in range(W)]
    counts = [hi - lo for lo, hi in sizes]
    if len(set(counts)) == 1:
        out = local.new_empty((total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(W)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def rank_seed(seed, rank, world_size=None):
    """Philox key of `rank` for the stream `seed` (the reference seeds `seed + local_rank`, runner/runner.py:39).
    The Philox counter of the sampling kernels is the LOCAL element index plus the timestep, so two shards with equal
    keys would draw identical noise.  `seed + rank` collides as soon as the caller's seeds are consecutive (batch index
    + 1 on rank r-1 == batch index on rank r); with `world_size` the key is `seed * world_size + rank`, which is unique
    per (seed, rank)."""
    if world_size is None:
        if dist.is_available() and dist.is_initialized():
            world_size = dist.get_world_size()
        else:
            world_size = 1
    assert 0 <= int(rank) < int(world_size)
    return int(seed) * int(world_size) + int(rank)
