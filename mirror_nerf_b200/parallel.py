"""Ray-parallel sharding across GPUs.  Rays are independent units -- a reflected ray is
spawned by and blended back into its own parent only -- so an image (or a ray batch) is split into contiguous shards of
whole 128-ray tiles, one process per GPU, weights replicated, and the render needs NO collective.  The only optional
exchange is gathering the compact per-ray outputs (rgb/depth, 16 B per ray) to every rank / rank 0; it goes through
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The reference's counterpart is a single line of PyTorch-Lightning config (R/train.py:582, DDP over ray batches).
"""
from __future__ import annotations

import torch

TILE = 128  # rays per scheduling unit (one tensor-core tile of the coarse pass is 2 rays x 64 samples; keep shards tile-aligned)


def shard_bounds(n: int, rank: int, world: int, tile: int = TILE):
    """[lo, hi) of rank's contiguous shard of n rays; shards are whole tiles, sizes differ by at most one tile."""
    tiles = (n + tile - 1) // tile
    per, rem = divmod(tiles, world)
    first = rank * per + min(rank, rem)
    count = per + (1 if rank < rem else 0)
    lo = min(first * tile, n)
    hi = min((first + count) * tile, n)
    return lo, hi


def shard_rays(rays: torch.Tensor, rank: int, world: int):
    lo, hi = shard_bounds(rays.shape[0], rank, world)
    return rays[lo:hi], (lo, hi)


def gather_rows(local: torch.Tensor, n: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Concatenate every rank's rows (shard_bounds order) on every rank.  `local` is this rank's (hi-lo, ...) slice."""
    import torch.distributed as dist
    if world == 1:
        return local
    bounds = [shard_bounds(n, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, bounds)], 0)


def render_sharded(render_fn, rays: torch.Tensor, rank: int, world: int, gather=("rgb_fine", "depth_fine"), group=None):
    """Render this rank's shard with `render_fn(rays_shard) -> dict` and gather the listed per-ray outputs.
    Returns (local_results, gathered dict or None)."""
    n = rays.shape[0]
    mine, _ = shard_rays(rays, rank, world)
    res = render_fn(mine)
    if not gather:
        return res, None
    return res, {k: gather_rows(res[k], n, rank, world, group) for k in gather if k in res}
