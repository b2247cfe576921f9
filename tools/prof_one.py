"""One eval-mode render_rays call (32768 rays, 64+128) repeated 3x -- the command profiled with ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_models  # noqa: E402
from mirror_nerf_b200.rendering import render_rays  # noqa: E402
from mirror_nerf_b200.synthetic import camera_rays  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "tc3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
models, emb = make_models()
rays = camera_rays(800, 800)[:: max(1, 640000 // n)][:n].contiguous().cuda()
with torch.no_grad():
    for _ in range(3):
        r = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False,
                        field_impl=impl)
torch.cuda.synchronize()
print("ok", float(r["rgb_fine"].mean()))
