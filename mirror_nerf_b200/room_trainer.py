"""Minimal trainer for the analytic mirror-room scene (room_scene.py): this repo's training path end to end -- render_rays under
autograd, the reference's one-bounce train-time recursion written with torch ops exactly as R/train.py:194-296 does it (ground-truth
mirror mask at level 0, secondary rays from x_surface / surface normal without detaching, blend by the mask), the loss terms of
R/losses.py that apply (colour on both passes, mirror-mask BCE, normal regularisation) and FlatDataParallel's Adam.
Used by tools/train_room.py (makes tests/golden/room_field.npz) and by the functional training test."""
from __future__ import annotations

import math

import torch

from .mirror_nerf import Embedding, MirrorNeRF
from .parallel import FlatDataParallel
from .rendering import render_rays
from .room_scene import random_room_rays, room_pose, trace_room
from .synthetic import camera_rays
from .trace import render_rays_recursive

NEAR, FAR = 0.05, 12.0
_EPS = torch.finfo(torch.float32).eps


def _l2n(x):
    return x / torch.sqrt(torch.clamp((x ** 2).sum(-1, keepdim=True), min=_EPS))


class RoomTrainer:
    def __init__(self, device="cuda", lr=5e-4, rays_per_step=4096, seed=0, peer_fused=False):
        torch.manual_seed(seed)
        self.dev = torch.device(device)
        self.models = {k: MirrorNeRF(predict_normal=True, predict_mirror_mask=True).to(self.dev).train() for k in ("coarse", "fine")}
        self.emb = {"xyz": Embedding(10), "dir": Embedding(4)}
        self.ddp = FlatDataParallel(self.models, lr=lr, peer_fused=peer_fused)
        self.base_lr, self.n = lr, rays_per_step
        self.gen = torch.Generator().manual_seed(seed + 1)
        self.args = (64, False, 1.0, 0.0, 128, 32768, False)

    def render_train(self, rays, mask_gt):
        r = render_rays(self.models, self.emb, rays, *self.args, test_time=False, compute_normal=False)
        m = mask_gt.bool()
        if bool(m.any()):
            n = _l2n(r["surface_normal_fine"])
            w = _l2n(-rays[:, 3:6])
            refl = 2 * (w * n).sum(-1, keepdim=True) * n - w
            sec = torch.cat([r["x_surface_fine"], refl, torch.full_like(rays[:, 6:7], 0.1), rays[:, 7:8]], -1)[m]
            child = render_rays(self.models, self.emb, sec, *self.args, test_time=False, compute_normal=False)
            for typ in ("coarse", "fine"):
                base = r[f"rgb_{typ}"]
                part = base.clone().detach()
                part[m] = child[f"rgb_{typ}"]
                m3 = mask_gt[:, None]
                r[f"rgb_{typ}"] = m3 * part + (1 - m3) * base
        return r

    def step(self, lr_scale=1.0):
        self.ddp.lr = self.base_lr * lr_scale
        rays = random_room_rays(self.n, self.gen, NEAR, FAR).to(self.dev)
        gt, mask_gt, _ = trace_room(rays)
        self.ddp.zero_grad()
        r = self.render_train(rays, mask_gt)
        loss = 0.0
        for typ in ("coarse", "fine"):
            loss = loss + ((r[f"rgb_{typ}"] - gt) ** 2).mean()
            mm = r[f"mirror_mask_{typ}"].clamp(1e-7, 1 - 1e-7)
            loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(mm, mask_gt)
            loss = loss + 1e-4 * (torch.relu(r[f"pred_normal_{typ}"] * rays[:, None, 3:6]).sum(-1) * r[f"weights_{typ}"]).mean()
        loss.backward()
        self.ddp.step()
        return loss.detach()

    def psnr(self, res=200, view=0):
        """(PSNR in dB against the analytic ground truth of a held-out view, predicted mirror fraction); eval path, one bounce."""
        rays = camera_rays(res, res, c2w=room_pose(view), near=NEAR, far=FAR).to(self.dev)
        gt, _, _ = trace_room(rays)
        with torch.no_grad():
            out = render_rays_recursive(self.models, self.emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
        mse = float(((out["rgb_fine"] - gt) ** 2).mean())
        return -10 * math.log10(mse), float((out["mirror_mask_fine"] != 0).float().mean())
