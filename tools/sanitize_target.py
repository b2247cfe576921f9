"""Small configurations of every round-2 kernel path -- target for compute-sanitizer (memcheck / racecheck / initcheck):
fused compositor (tc3, tc2, ragged sample counts), device-side recursion with the roughness cone (slabs), analytic-normal kernel,
both schedules of the tc2 kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import make_models
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.synthetic import random_rays
from mirror_nerf_b200.trace import render_rays_recursive
models, emb = make_models()
rays = random_rays(301, seed=5).cuda()
with torch.no_grad():
    for impl in ("tc3", "tc2", "tc1"):
        render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False, field_impl=impl)
        render_rays(models, emb, rays[:37], 40, True, 0, 0, 33, 32768, True, test_time=True, compute_normal=False, field_impl=impl)
    render_rays(models, emb, rays[:64], 64, False, 0, 0, 128, 32768, False, test_time=False, compute_normal=True)
    r = render_rays_recursive(models, emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=2, compact_outputs=True,
                              normal_noise_std=0.05, trace_ray_times=2, workspace_budget_bytes=64 << 20, field_impl="tc2")
    # the N-split schedule of the tc2 kernels (fused and per-point), full and sigma-only launches
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    lib.mnrf_debug_set_tc_schedule(1)
    for fused in (True, False):
        render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False, field_impl="tc2",
                    fused_composite=fused)
        render_rays(models, emb, rays[:37], 40, True, 0, 0, 33, 32768, True, test_time=False, compute_normal=False, field_impl="tc2",
                    fused_composite=fused)
    lib.mnrf_debug_set_tc_schedule(-1)
torch.cuda.synchronize()
print("ok", float(r["rgb_fine"].sum()))
