"""Analytic synthetic scene (no dataset is available offline): a closed box room [-3,3]^3 with procedurally
textured walls and a planar mirror on the x = +3 wall, ray-traced exactly (one reflection).  Used to make a *scene-like* field
(tools/train_room.py trains the renderer on it) and to give bench.py / the tests a ground truth for PSNR.  Pure torch elementwise
ops on whatever device the rays live on -- synthetic data generation, not part of the render path."""
from __future__ import annotations

import math

import torch

HALF = 3.0
MIRROR_HALF = 1.0  # mirror: x = +3, |y| <= 1, |z| <= 1
_BASE = torch.tensor([[0.85, 0.35, 0.30], [0.30, 0.75, 0.40],    # -x, +x
                      [0.30, 0.45, 0.85], [0.85, 0.80, 0.35],    # -y, +y
                      [0.70, 0.40, 0.80], [0.40, 0.80, 0.80]])   # -z, +z


def _hit_box(o, d):
    """First exit of rays starting inside the box: (t, wall index 0..5 = -x,+x,-y,+y,-z,+z)."""
    eps = 1e-9
    par = d.abs() < eps                                    # parallel to the slab: never exits through it
    inv = 1.0 / torch.where(par, torch.ones_like(d), d)
    t_axis = torch.where(d > 0, (HALF - o) * inv, (-HALF - o) * inv)   # exit distance per axis
    t_axis = torch.where(par, torch.full_like(d, float("inf")), t_axis)
    t, axis = t_axis.min(dim=-1)
    sign_pos = torch.gather(d, 1, axis[:, None])[:, 0] > 0
    wall = axis * 2 + sign_pos.long()
    return t, wall


def _wall_color(p, wall):
    """Base colour per wall modulated by a unit checker and a soft radial shade (textured, view independent)."""
    base = _BASE.to(p.device)[wall]
    axis = wall // 2
    idx = torch.stack([(axis + 1) % 3, (axis + 2) % 3], -1)
    uv = torch.gather(p, 1, idx)
    checker = ((torch.floor(uv[:, 0]) + torch.floor(uv[:, 1])) % 2 == 0).float()
    shade = 0.75 + 0.25 * torch.cos(0.8 * uv[:, 0]) * torch.cos(0.8 * uv[:, 1])
    return base * (0.55 + 0.45 * checker)[:, None] * shade[:, None]


def trace_room(rays):
    """rays (n,8) [o, d, near, far] with origins inside the room.  Returns rgb (n,3), mirror_mask (n,) in {0,1}, depth (n,)."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    d = d / d.norm(dim=-1, keepdim=True)
    t, wall = _hit_box(o, d)
    p = o + d * t[:, None]
    mirror = (wall == 1) & (p[:, 1].abs() <= MIRROR_HALF) & (p[:, 2].abs() <= MIRROR_HALF)
    rgb = _wall_color(p, wall)
    if bool(mirror.any()):
        n = torch.tensor([-1.0, 0.0, 0.0], device=rays.device)
        dm = d[mirror]
        r = dm - 2 * (dm @ n)[:, None] * n
        om = p[mirror] + 1e-4 * n
        t2, wall2 = _hit_box(om, r)
        rgb = rgb.clone()
        rgb[mirror] = 0.9 * _wall_color(om + r * t2[:, None], wall2) + 0.05
    return rgb, mirror.float(), t


def room_pose(i=0):
    """Camera inside the room looking at the mirror wall (+x); small orbit of poses indexed by i (3x4 camera-to-world)."""
    a = 0.25 * math.sin(0.9 * i)                      # yaw around +y
    pos = torch.tensor([-1.6 + 0.3 * math.cos(0.7 * i), 0.2 * math.sin(1.3 * i), 0.8 * math.sin(0.5 * i)])
    fwd = torch.tensor([math.cos(a), 0.0, math.sin(a)])   # viewing direction
    up = torch.tensor([0.0, 1.0, 0.0])
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    up2 = torch.linalg.cross(right, fwd)
    # camera looks along -z of its own frame (R/datasets/ray_utils.py:6-30): columns = [right, up, -forward]
    R = torch.stack([right, up2, -fwd], 1)
    return torch.cat([R, pos[:, None]], 1)


def random_room_rays(n, generator=None, near=0.05, far=12.0, device="cpu"):
    """Training rays: origins in the middle of the room, directions biased toward the mirror wall (half of them) or uniform."""
    g = generator
    o = torch.rand(n, 3, generator=g) * torch.tensor([2.5, 3.0, 3.0]) + torch.tensor([-2.5, -1.5, -1.5])
    d = torch.randn(n, 3, generator=g)
    toward = torch.rand(n, generator=g) < 0.5
    tgt = torch.stack([torch.full((n,), HALF), (torch.rand(n, generator=g) * 2 - 1) * 2.5, (torch.rand(n, generator=g) * 2 - 1) * 2.5], 1)
    d = torch.where(toward[:, None], tgt - o, d)
    d = d / d.norm(dim=-1, keepdim=True)
    nf = torch.tensor([near, far]).expand(n, 2)
    return torch.cat([o, d, nf], 1).contiguous().to(device)
