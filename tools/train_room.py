"""Train the renderer on the analytic room scene (mirror_nerf_b200/room_scene.py) with the training path of this repo
(render_rays under autograd, one-bounce recursion as R/train.py:194-296 does it with torch ops, FlatDataParallel Adam) and save
the resulting *scene-like* field as tests/golden/room_field.npz (fp16 storage; loaded back as fp32).

    python tools/train_room.py [--steps 6000] [--rays 4096] [--out tests/golden/room_field.npz]

Prints the PSNR against the analytic ground truth on a held-out view as training proceeds."""
import argparse, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
from mirror_nerf_b200.parallel import FlatDataParallel
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.room_scene import random_room_rays, room_pose, trace_room
from mirror_nerf_b200.synthetic import camera_rays
from mirror_nerf_b200.trace import render_rays_recursive

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6000)
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--lr", type=float, default=5e-4)
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "room_field.npz"))
args = ap.parse_args()

torch.manual_seed(0)
dev = torch.device("cuda")
models = {k: MirrorNeRF(predict_normal=True, predict_mirror_mask=True).to(dev).train() for k in ("coarse", "fine")}
emb = {"xyz": Embedding(10), "dir": Embedding(4)}
ddp = FlatDataParallel(models, lr=args.lr)
NEAR, FAR = 0.05, 12.0
rargs = (64, False, 1.0, 0.0, 128, 32768, False)
eps = torch.finfo(torch.float32).eps


def l2n(x):
    return x / torch.sqrt(torch.clamp((x ** 2).sum(-1, keepdim=True), min=eps))


def render_train(rays, mask_gt):
    """Level 0 + one bounce for the mirror rays (train semantics: ground-truth mask at level 0, R/train.py:155-166,248-252)."""
    r = render_rays(models, emb, rays, *rargs, test_time=False, compute_normal=False)
    m = mask_gt.bool()
    if bool(m.any()):
        n = l2n(r["surface_normal_fine"])
        w = l2n(-rays[:, 3:6])
        refl = 2 * (w * n).sum(-1, keepdim=True) * n - w
        sec = torch.cat([r["x_surface_fine"], refl, torch.full_like(rays[:, 6:7], 0.1), rays[:, 7:8]], -1)[m]
        child = render_rays(models, emb, sec, *rargs, test_time=False, compute_normal=False)
        for typ in ("coarse", "fine"):
            base = r[f"rgb_{typ}"]
            part = base.clone().detach()
            part[m] = child[f"rgb_{typ}"]
            m3 = mask_gt[:, None]
            r[f"rgb_{typ}"] = m3 * part + (1 - m3) * base
    return r


def psnr_eval():
    rays = camera_rays(200, 200, c2w=room_pose(0), near=NEAR, far=FAR).to(dev)
    gt, _, _ = trace_room(rays)
    with torch.no_grad():
        out = render_rays_recursive(models, emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
    mse = float(((out["rgb_fine"] - gt) ** 2).mean())
    return -10 * math.log10(mse), float((out["mirror_mask_fine"] != 0).float().mean())


g = torch.Generator().manual_seed(1)
t0 = time.time()
for it in range(args.steps + 1):
    if it % 500 == 0:
        p, mf = psnr_eval()
        print(f"step {it:5d}  psnr {p:6.2f} dB  predicted mirror fraction {mf:.3f}  ({time.time() - t0:.0f} s)", flush=True)
    if it == args.steps:
        break
    ddp.lr = args.lr * (1.0 if it < args.steps // 2 else (0.4 if it < 5 * args.steps // 6 else 0.15))  # step decay (R/opt.py: steplr)
    rays = random_room_rays(args.rays, g, NEAR, FAR).to(dev)
    gt, mask_gt, _ = trace_room(rays)
    ddp.zero_grad()
    r = render_train(rays, mask_gt)
    loss = 0.0
    for typ in ("coarse", "fine"):
        loss = loss + ((r[f"rgb_{typ}"] - gt) ** 2).mean()
        mm = r[f"mirror_mask_{typ}"].clamp(1e-7, 1 - 1e-7)
        loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(mm, mask_gt)
        loss = loss + 1e-4 * (torch.relu(r[f"pred_normal_{typ}"] * rays[:, None, 3:6]).sum(-1) * r[f"weights_{typ}"]).mean()
    loss.backward()
    ddp.step()
    if it % 100 == 0:
        print(f"  it {it} loss {float(loss.detach()):.5f}", flush=True)

sd = {}
for k, m in models.items():
    for name, p in m.state_dict().items():
        sd[f"{k}/{name}"] = p.detach().cpu().numpy().astype(np.float16)
np.savez_compressed(args.out, **sd)
print("saved", args.out, os.path.getsize(args.out) // 1024, "KiB")
