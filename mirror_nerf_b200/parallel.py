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

    def __init__(self, models, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None, peer_fused=False):
        """peer_fused=True (NCCL world > 1 on one NVLink/NVSwitch node): the flat buffers live in symmetric memory
        (torch.distributed._symmetric_memory) and step() runs ONE kernel, mnrf_peer_allreduce_adam -- reduce-scatter of the
        gradients with P2P loads, Adam on the owned shard (sharded moment estimates), all-gather of the new parameters with P2P
        stores -- between two device-side barriers, instead of ncclAllReduce + mnrf_adam_step."""
        mods = list(models.values()) if isinstance(models, dict) else list(models)
        self.modules = mods
        self.params = [p for m in mods for p in m.parameters()]  # R/utils/__init__.py:33-45 order
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if dt != torch.float32 or any(p.dtype != dt or p.device != dev for p in self.params):
            raise RuntimeError("FlatDataParallel: all parameters must be float32 on one device")
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.peer = None
        if peer_fused:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and dev.type == "cuda"):
                raise RuntimeError("peer_fused=True needs an initialised NCCL process group with world_size > 1 on CUDA")
            pg = group if group is not None else dist.group.WORLD
            n_pad = (n + 3) // 4 * 4
            self.flat_params = symm_mem.empty(n_pad, dtype=dt, device=dev)
            self.flat_grads = symm_mem.empty(n_pad, dtype=dt, device=dev)
            self.flat_params.zero_()
            self.flat_grads.zero_()
            hp = symm_mem.rendezvous(self.flat_params, pg.group_name)
            hg = symm_mem.rendezvous(self.flat_grads, pg.group_name)
            self.peer = dict(hp=hp, hg=hg, world=hp.world_size, rank=hp.rank, n_pad=n_pad)
        else:
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
        if self.peer is not None:
            import ctypes as C
            from . import _lib
            lo, hi = C.c_int64(), C.c_int64()
            _lib.check(_lib.load().mnrf_peer_shard(self.peer["n_pad"], self.peer["world"], self.peer["rank"], C.byref(lo), C.byref(hi)))
            self.peer["shard"] = (lo.value, hi.value)
            self.exp_avg = torch.zeros(max(hi.value - lo.value, 1), device=dev, dtype=dt)     # moments of the owned shard only
            self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        else:
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
        if self.peer is not None:
            return self._step_peer()
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

    def _step_peer(self):
        """reduce-scatter + Adam + all-gather as one kernel over peer memory, fenced by two device-side barriers."""
        import ctypes as C
        from . import _lib
        from .mirror_nerf import invalidate_packed
        pr = self.peer
        self.step_count += 1
        lib = _lib.load()
        W = pr["world"]
        gp = (C.c_uint64 * W)(*[int(x) for x in pr["hg"].buffer_ptrs])
        pp = (C.c_uint64 * W)(*[int(x) for x in pr["hp"].buffer_ptrs])
        with torch.cuda.device(self.flat_params.device):
            pr["hg"].barrier(channel=0)   # every rank's backward has written its gradients
            _lib.check(lib.mnrf_peer_allreduce_adam(
                gp, pp, W, pr["rank"], C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()), pr["n_pad"],
                self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mnrf_peer_allreduce_adam")
            pr["hp"].barrier(channel=1)   # every owner's parameter shard has landed in this rank's buffer
        for m in self.modules:
            invalidate_packed(m)
