"""Ray generation on the device: what R/datasets/ray_utils.py::get_ray_directions + get_rays and
R/datasets/blender.py:158-168 do on the CPU, as one CUDA kernel, so a frame's (H*W, 8) ray tensor never crosses PCIe."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .mirror_nerf import _ptr, _stream_ptr


def focal_from_fov(W: int, camera_angle_x: float) -> float:
    """blender.py:33-35: focal = 0.5 * W / tan(0.5 * camera_angle_x)."""
    return 0.5 * W / math.tan(0.5 * camera_angle_x)


def generate_rays(H: int, W: int, focal: float, c2w, near: float, far: float, device="cuda") -> torch.Tensor:
    """(H*W, 8) = [rays_o, rays_d (normalised), near, far] on `device`; c2w: (3,4) camera-to-world (any array-like)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32, device="cpu").reshape(-1)
    if c2w.numel() not in (12, 16):
        raise RuntimeError("c2w must be 3x4 or 4x4")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("generate_rays: CUDA device required (no CPU path)")
    rays = torch.empty(H * W, 8, device=dev, dtype=torch.float32)
    arr = (C.c_float * 12)(*[float(v) for v in c2w[:12]])
    with torch.cuda.device(dev):
        lib = _lib.load()
        _lib.check(lib.mnrf_generate_rays(H, W, float(focal), arr, float(near), float(far), _ptr(rays), _stream_ptr()),
                   "mnrf_generate_rays")
    return rays
