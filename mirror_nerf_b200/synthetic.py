"""Deterministic synthetic inputs (no dataset / checkpoint is available offline).

* ``make_state_dict``  random-init MirrorNeRF weights with the reference's state_dict keys and
  [out,in] fp32 layout (R/models/mirror_nerf.py:60-99).  numpy PCG64 so that the values are the same
  on every machine; the sigma head can be scaled so rays saturate (SURVEY.md section 8d "adversarial").
* ``random_rays``      o ~ U[-1,1]^3, d = normalize(N(0,I)), near/far constants.
* ``camera_rays``      pinhole camera exactly as R/datasets/ray_utils.py:6-53 + R/datasets/blender.py:158-168.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def field_param_shapes(D=8, W=256, in_xyz=63, in_dir=27, skips=(4,), predict_normal=True,
                       predict_mirror_mask=True):
    s = OrderedDict()
    for i in range(D):
        fan_in = in_xyz if i == 0 else (W + in_xyz if i in skips else W)
        s[f"xyz_encoding_{i + 1}.0.weight"] = (W, fan_in)
        s[f"xyz_encoding_{i + 1}.0.bias"] = (W,)
    s["xyz_encoding_final.weight"] = (W, W)
    s["xyz_encoding_final.bias"] = (W,)
    s["dir_encoding.0.weight"] = (W // 2, W + in_dir)
    s["dir_encoding.0.bias"] = (W // 2,)
    s["sigma.weight"] = (1, W)
    s["sigma.bias"] = (1,)
    s["rgb.0.weight"] = (3, W // 2)
    s["rgb.0.bias"] = (3,)
    if predict_normal:
        s["normal_net.0.weight"] = (W // 2, W)
        s["normal_net.0.bias"] = (W // 2,)
        s["normal_net.1.weight"] = (3, W // 2)
        s["normal_net.1.bias"] = (3,)
    if predict_mirror_mask:
        s["is_mirror_net.0.weight"] = (W // 2, W)
        s["is_mirror_net.0.bias"] = (W // 2,)
        s["is_mirror_net.2.weight"] = (1, W // 2)
        s["is_mirror_net.2.bias"] = (1,)
    return s


def make_state_dict(seed=0, sigma_scale=40.0, sigma_bias=None, predict_normal=True,
                    predict_mirror_mask=True, device="cpu", mirror_scale=1.0):
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init from numpy PCG64(seed)."""
    g = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for k, shp in field_param_shapes(predict_normal=predict_normal,
                                     predict_mirror_mask=predict_mirror_mask).items():
        if k.endswith("weight"):
            fan_in = shp[1]
        else:
            fan_in = field_param_shapes(predict_normal=predict_normal,
                                        predict_mirror_mask=predict_mirror_mask)[k[:-4] + "weight"][1]
        b = 1.0 / math.sqrt(fan_in)
        sd[k] = torch.from_numpy(g.uniform(-b, b, size=shp).astype(np.float32))
    sd["sigma.weight"] = sd["sigma.weight"] * float(sigma_scale)
    if sigma_bias is not None:
        sd["sigma.bias"] = torch.full((1,), float(sigma_bias), dtype=torch.float32)
    if predict_mirror_mask and mirror_scale != 1.0:
        sd["is_mirror_net.2.weight"] = sd["is_mirror_net.2.weight"] * float(mirror_scale)
    return OrderedDict((k, v.to(device)) for k, v in sd.items())


# The synthetic "scene" every fixture, test, smoke() and bench.py uses: coarse = seed 0, fine = seed 4 (about half of
# random rays saturate, the rest stay translucent), sigma head x40 (sharp, adversarial density: SURVEY.md 7.3), mirror
# head x100 so that the composited mirror probability straddles the 0.5 threshold (about half of the rays bounce).
SCENE_SEEDS = {"coarse": 0, "fine": 4}
SCENE_MIRROR_SCALE = 100.0


def scene_state_dicts(predict_normal=True, predict_mirror_mask=True, device="cpu"):
    return {k: make_state_dict(seed, 40.0, None, predict_normal, predict_mirror_mask, device,
                               mirror_scale=SCENE_MIRROR_SCALE)
            for k, seed in SCENE_SEEDS.items()}


def random_rays(n, seed=1, near=0.05, far=8.0, device="cpu"):
    g = np.random.Generator(np.random.PCG64(seed))
    o = g.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    d = g.standard_normal(size=(n, 3)).astype(np.float32)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    nf = np.empty((n, 2), np.float32)
    nf[:, 0] = near
    nf[:, 1] = far
    return torch.from_numpy(np.concatenate([o, d.astype(np.float32), nf], 1)).to(device)


def camera_rays(H=800, W=800, fov_x=0.6911112070083618, c2w=None, near=0.05, far=8.0, device="cpu"):
    """(H*W, 8) rays of one pinhole view.  Directions [(i-W/2)/f, -(j-H/2)/f, -1] (no +0.5), rotated by
    c2w[:3,:3] and normalised; origin c2w[:3,3] (R/datasets/ray_utils.py:6-53)."""
    focal = 0.5 * W / math.tan(0.5 * fov_x)
    if c2w is None:
        c2w = torch.tensor([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 2.5]])
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                          indexing="ij")
    dirs = torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)
    rd = dirs.reshape(-1, 3) @ c2w[:3, :3].T
    rd = rd / torch.norm(rd, dim=-1, keepdim=True)
    ro = c2w[:3, 3].expand(rd.shape)
    nf = torch.tensor([near, far], dtype=torch.float32).expand(rd.shape[0], 2)
    return torch.cat([ro, rd, nf], 1).contiguous().to(device)
