"""One tiny training step (6 rays, 64+128 samples, analytic normals, ray gradients) -- target for compute-sanitizer runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_train import _loss, _models, _rng, _smooth_sds
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.synthetic import random_rays
models, emb = _models(_smooth_sds())
rays = random_rays(6, seed=3).cuda().requires_grad_(True)
r = render_rays(models, emb, rays, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False, compute_normal=True, rng=_rng(6))
_loss(r, rays[:, 3:6].detach(), 0).backward()
torch.cuda.synchronize()
print("tiny train step ok", float(rays.grad.abs().sum()))
