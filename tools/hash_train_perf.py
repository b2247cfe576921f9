"""Timing of the hash-grid training pass (csrc/train_hash.cu) on one GPU: forward / backward of the coarse (64) and fine (192)
passes of a 4096-ray batch, and the whole optimisation step as bench.py's hash_grid_train_step measures it."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    from mirror_nerf_b200.mirror_nerf import Embedding
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import camera_rays
    out = {}
    models = bench.hash_models(dev, sigma_scale=5.0)
    for m in models.values():
        m.train()
    emb = {"xyz": Embedding(0), "dir": Embedding(0)}
    c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.9]])
    allrays = camera_rays(800, 800, c2w=c2w, near=0.05, far=2.0)
    rays = allrays[torch.randperm(allrays.shape[0], generator=torch.Generator().manual_seed(1))[:4096]].contiguous().to(dev)

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for cn in (True, False):
        state = {}

        def fwd():
            state["r"] = render_rays(models, emb, rays, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False, compute_normal=cn)

        def fwd_bwd():
            for m in models.values():
                m.zero_grad(set_to_none=True)
            fwd()
            r = state["r"]
            loss = (r["rgb_fine"] ** 2).mean() + (r["rgb_coarse"] ** 2).mean() + r["mirror_mask_fine"].mean()
            if cn:
                loss = loss + 1e-3 * r["normal_dif_fine"].mean() + 1e-3 * r["normal_dif_coarse"].mean()
            loss.backward()

        out[f"forward_ms_compute_normal_{int(cn)}"] = timed(fwd)
        out[f"forward_backward_ms_compute_normal_{int(cn)}"] = timed(fwd_bwd)
    del models
    out["step"] = bench.hash_train_bench(dev, 3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
