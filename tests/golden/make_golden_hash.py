"""Regression vectors of the hash-grid oracle (oracle/hashgrid_oracle.py) -> tests/golden/hash_field.npz.

    python tests/golden/make_golden_hash.py

NOT reference output: the reference's encoder is tinycudann, which is neither vendored nor runnable in the build container, so
the hash-grid path stays PARITY UNPINNED (see the oracle's header).  These vectors only pin the restatement itself -- level table,
table indices (through the encoded features), field outputs, analytic normals and training gradients on seeded inputs -- so
that a later edit of the oracle cannot silently move the target the CUDA kernels are tested against.  The 48.8 MB table is
regenerated from its seed (numpy PCG64), only the outputs are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def generate():
    from oracle import hashgrid_oracle as H
    torch.set_num_threads(8)
    out = {}
    for bound in (1.0, 2.0):
        tag = f"b{int(bound)}"
        lv, total = H.level_table(bound)
        out[f"levels_{tag}"] = np.array([[s, r, o, n] for s, r, o, n in lv], np.float64)
        out[f"total_{tag}"] = np.array([total], np.int64)
        sd = H.make_state_dict(3, bound=bound)
        g = torch.Generator().manual_seed(1)
        n = 96
        xyz = (torch.rand(n, 3, generator=g) * 2 - 1) * bound * 1.05
        xyz[0], xyz[1], xyz[2] = 0.0, bound, -bound
        d = torch.randn(n, 3, generator=g)
        d = d / d.norm(dim=-1, keepdim=True)
        x = torch.cat([xyz, d], 1)
        out[f"x_{tag}"] = x.numpy()
        out[f"enc_{tag}"] = H.hashgrid_encode(sd["encoder.params"], (xyz + bound) / (2 * bound), bound).numpy()
        o = H.field_forward(sd, x, bound=bound, compute_normal=True)
        for k in ("sigma", "rgb", "pred_normal", "is_mirror", "normal"):
            out[f"{k}_{tag}"] = o[k].detach().numpy()
    # training gradients of a tiny weighted loss (incl. the double backward through the analytic normal)
    sd = H.make_state_dict(3, sigma_scale=4.0)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    g = torch.Generator().manual_seed(2)
    x = torch.cat([torch.rand(40, 3, generator=g) * 1.9 - 0.95, torch.nn.functional.normalize(torch.randn(40, 3, generator=g), dim=-1)], 1)
    c = torch.randn(40, 11, generator=g)
    o = H.field_forward(p, x, compute_normal=True)
    loss = (o["sigma"][:, 0] * c[:, 0]).sum() + (o["rgb"] * c[:, 1:4]).sum() + (o["is_mirror"][:, 0] * c[:, 4]).sum() + \
        (o["pred_normal"] * c[:, 5:8]).sum() + (o["normal"] * c[:, 8:11]).sum()
    loss.backward()
    out["train_x"], out["train_c"] = x.numpy(), c.numpy()
    for k, v in p.items():
        if k == "encoder.params":
            nz = torch.nonzero(v.grad).flatten()
            out["grad_table_idx"], out["grad_table_val"] = nz.numpy(), v.grad[nz].numpy()
        else:
            out["grad_" + k] = v.grad.numpy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "hash_field.npz"), **generate())
    print("wrote", os.path.join(HERE, "hash_field.npz"))
