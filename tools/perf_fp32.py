import sys, time, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from util import make_models
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.synthetic import random_rays
models, emb = make_models()
n=4096
rays = random_rays(n, seed=4).cuda()
for cn, impl in ((True,'fp32'),(True,'tc3'),(True,'tc1'),(False,'tc3')):
    with torch.no_grad():
        for _ in range(2): render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=False, compute_normal=cn, field_impl=impl)
        torch.cuda.synchronize(); t0=time.time()
        for _ in range(3): render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=False, compute_normal=cn, field_impl=impl)
        torch.cuda.synchronize(); dt=(time.time()-t0)/3
    print(f"train-mode forward (both passes full) compute_normal={cn} impl={impl}: {n/dt:,.0f} rays/s ({dt*1e3:.1f} ms / {n} rays)")
