"""Throughput of one eval render level (64+128 samples) with the hash-grid field (BASELINE config 3) on an 800x800 view."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mirror_nerf_b200.mirror_nerf import Embedding
from mirror_nerf_b200.mirror_nerf_tcnn import MirrorNeRFTcnn
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.trace import render_rays_recursive
from mirror_nerf_b200.synthetic import camera_rays
from oracle import hashgrid_oracle as H  # synthetic weights only
models = {}
for k, seed in (("coarse", 7), ("fine", 8)):
    m = MirrorNeRFTcnn(bound=1, predict_normal=True, predict_mirror_mask=True); m.load_state_dict(H.make_state_dict(seed)); models[k] = m.cuda().eval()
emb = {"xyz": Embedding(0), "dir": Embedding(0)}
c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.9]])
rays = camera_rays(800, 800, c2w=c2w, near=0.05, far=2.0).cuda()
n = rays.shape[0]
def level(): return render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
def bounce(): return render_rays_recursive(models, emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
with torch.no_grad():
    for fn, name, levels in ((level, "one level", 1), (bounce, "1 bounce (eval semantics)", 2)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): r = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        pts = n * 256 * levels
        print(f"hash-grid field, {name}: {ms:.1f} ms/image  {n / ms * 1e3:.0f} rays/s  {pts / ms / 1e6:.2f} Gpoints/s  table reads {pts * 128 * 8 / ms / 1e9:.2f} TB/s (L2+HBM, algorithmic 1 KB/point)")
print("mirror fraction", float((r["mirror_mask_fine"] != 0).float().mean()))
