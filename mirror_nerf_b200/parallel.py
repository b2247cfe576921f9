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


# ------------------------------------------------------------------------------------------------------------------
# Data-parallel training (BASELINE config 5): what PL DDP + torch.optim.Adam do around the hot path in
# R/train.py:368-375,582 and R/utils/__init__.py:47-58 -- ONE sum all-reduce of a flat fp32 gradient buffer per step
# (2 x 662,152 elements = 5.3 MB for coarse + fine) and one Adam kernel over the flat parameter buffer.
# ------------------------------------------------------------------------------------------------------------------
class FlatDataParallel:
    """Flat parameter / gradient buffers for a dict (or list) of models.

    Every parameter becomes a view into ``flat_params`` and gets a ``.grad`` view into ``flat_grads`` (autograd
    accumulates into it in place), so that the gradient exchange is a single ``all_reduce`` and the optimizer a single
    kernel launch (``mnrf_adam_step``; the 1/world average is folded into it).  ``group`` = a torch.distributed process
    group (NCCL on GPUs; gloo in the CPU tests, where only the exchange is exercised -- the Adam kernel needs a GPU).
    """

    def __init__(self, models, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None):
        mods = list(models.values()) if isinstance(models, dict) else list(models)
        self.modules = mods
        self.params = [p for m in mods for p in m.parameters()]  # R/utils/__init__.py:33-45 order
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if dt != torch.float32 or any(p.dtype != dt or p.device != dev for p in self.params):
            raise RuntimeError("FlatDataParallel: all parameters must be float32 on one device")
        n = sum(p.numel() for p in self.params)
        self.flat_params = torch.empty(n, device=dev, dtype=dt)
        self.flat_grads = torch.zeros(n, device=dev, dtype=dt)
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_params[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat_params[o:o + k].view(p.shape)
                p.grad = self.flat_grads[o:o + k].view(p.shape)
                o += k
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.group = group
        self.step_count = 0

    @property
    def world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def zero_grad(self):
        self.flat_grads.zero_()
        o = 0
        for p in self.params:  # a caller may have reset .grad (zero_grad(set_to_none=True)): re-attach the views
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grads.data_ptr() + 4 * o:
                p.grad = self.flat_grads[o:o + k].view(p.shape)
            o += k

    def all_reduce_grads(self):
        """SUM over ranks of the flat gradient buffer (the averaging happens in step())."""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat_grads

    def step(self):
        """all-reduce + Adam on the flat buffers; invalidates the packed kernel weights of the modules."""
        import ctypes as C
        from . import _lib
        from .mirror_nerf import invalidate_packed
        if not self.flat_params.is_cuda:
            raise RuntimeError("FlatDataParallel.step: the optimizer kernel runs on CUDA only (no CPU path)")
        self.all_reduce_grads()
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.flat_params.device):
            _lib.check(lib.mnrf_adam_step(
                C.c_void_p(self.flat_params.data_ptr()), C.c_void_p(self.flat_grads.data_ptr()),
                C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()), self.flat_params.numel(),
                self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0 / self.world,
                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mnrf_adam_step")
        for m in self.modules:
            invalidate_packed(m)
