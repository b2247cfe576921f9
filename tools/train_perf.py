"""Time one training step (render_rays forward + backward, 64+128 samples, compute_normal) on N rays."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.synthetic import random_rays, scene_state_dicts
from mirror_nerf_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
models = {}
for k, sd in scene_state_dicts().items():
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True); m.load_state_dict(sd); models[k] = m.cuda()
emb = {"xyz": Embedding(10), "dir": Embedding(4)}
rays = random_rays(n, seed=1).cuda()
tgt = torch.rand(n, 3, device="cuda")
def step():
    r = render_rays(models, emb, rays, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False, compute_normal=True)
    loss = sum(((r[f"rgb_{t}"] - tgt) ** 2).mean() + 1e-4 * r[f"normal_dif_{t}"].mean() + r[f"mirror_mask_{t}"].mean() for t in ("coarse", "fine"))
    for m in models.values():
        m.zero_grad(set_to_none=True)
    loss.backward()
    return loss
for i in range(2): step()
torch.cuda.synchronize()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
t0 = time.perf_counter()
e0.record()
for i in range(K): step()
e1.record()
t_host = (time.perf_counter() - t0) / K * 1e3
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
flops = n * 256 * 6.9e6
print(f"host enqueue time per step: {t_host:.1f} ms")
print(f"train step {n} rays: {ms:.1f} ms  {n / ms * 1e3:.0f} rays/s  ~{flops / ms / 1e9:.1f} TFLOP/s  launches/step {(_lib.launch_count() - l0) / K:.0f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
