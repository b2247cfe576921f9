"""Data-parallel gradient parity on real GPUs (run under torchrun with N ranks): the flat NCCL all-reduce of FlatDataParallel
over N half-batches, averaged, equals the single-process gradient of the mean loss over the whole batch (what PL DDP guarantees,
R/train.py:582).  Prints one line per rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
from mirror_nerf_b200.parallel import FlatDataParallel
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.room_scene import random_room_rays, trace_room
from mirror_nerf_b200.synthetic import make_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
emb = {"xyz": Embedding(10), "dir": Embedding(4)}
n = 256 * world
g = torch.Generator().manual_seed(0)
rays_all = random_room_rays(n, g)
gt_all, _, _ = trace_room(rays_all)
rng = {"perturb_u": torch.rand(n, 64, generator=g), "u_pdf": torch.rand(n, 128, generator=g)}


def build():
    models = {}
    for k, seed in (("coarse", 21), ("fine", 22)):
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(make_state_dict(seed, 5.0))
        models[k] = m.to(dev).train()
    return models, FlatDataParallel(models)


def grads(models, ddp, lo, hi):
    ddp.zero_grad()
    sl = lambda t: t[lo:hi].to(dev)
    r = render_rays(models, emb, sl(rays_all), 64, False, 1.0, 0.0, 128, 32768, False, test_time=False, compute_normal=True,
                    rng={k: v[lo:hi] for k, v in rng.items()})
    loss = sum(((r[f"rgb_{t}"] - sl(gt_all)) ** 2).mean() + 1e-2 * r[f"normal_dif_{t}"].mean() for t in ("coarse", "fine"))
    loss.backward()
    return float(loss.detach())


models, ddp = build()
per = n // world
l_mine = grads(models, ddp, rank * per, (rank + 1) * per)
ddp.all_reduce_grads()
avg = ddp.flat_grads.double() / world
if rank == 0:
    m1, d1 = build()
    d1.group = None
    l_full = grads(m1, d1, 0, n)                      # whole batch on one GPU (no collective)
    ref = d1.flat_grads.double()
    cos = float((avg * ref).sum() / (avg.norm() * ref.norm()))
    rel = float((avg - ref).norm() / ref.norm())
    print(f"ddp_check world={world}: cosine(all-reduced mean of {world} x {per}-ray shards, {n}-ray single-GPU gradient) = {cos:.8f}, "
          f"relative difference {rel:.2e}, elements {ref.numel()}", flush=True)
    assert cos > 0.99999 and rel < 5e-3
dist.barrier()
dist.destroy_process_group()
