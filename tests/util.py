"""Shared helpers for the parity tests."""
import numpy as np
import torch

from mirror_nerf_b200.synthetic import make_state_dict


def T(x, device="cpu"):
    return torch.from_numpy(np.asarray(x)).to(device)


def make_models(device="cuda", seeds=(0, 1), sigma_scale=40.0, predict_normal=True, predict_mirror_mask=True):
    """Two of OUR MirrorNeRF modules loaded with the synthetic state dicts the golden vectors were made with."""
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    models = {}
    for name, seed in zip(("coarse", "fine"), seeds):
        m = MirrorNeRF(predict_normal=predict_normal, predict_mirror_mask=predict_mirror_mask)
        m.load_state_dict(make_state_dict(seed, sigma_scale, predict_normal=predict_normal,
                                          predict_mirror_mask=predict_mirror_mask))
        models[name] = m.to(device).eval()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    return models, emb


def err_stats(got, want, floor=None):
    """Relative error |got-want| / max(|want|, floor) as (median, p99, max, frac>1e-3).  `floor` defaults to the
    RMS of `want` so that near-zero entries are judged on the tensor's own scale."""
    got = got.detach().double().cpu().flatten()
    want = want.detach().double().cpu().flatten()
    if floor is None:
        floor = max(float(want.pow(2).mean().sqrt()), 1e-12)
    e = (got - want).abs() / want.abs().clamp_min(floor)
    if e.numel() == 0:
        return dict(median=0.0, p99=0.0, max=0.0, frac=0.0)
    q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.double)) if e.numel() < 1_000_000 else \
        torch.tensor([e.median(), e.kthvalue(int(0.99 * e.numel()))[0]])
    return dict(median=float(q[0]), p99=float(q[1]), max=float(e.max()), frac=float((e > 1e-3).double().mean()))


def fmt_stats(name, s):
    return f"{name:28s} median {s['median']:.2e}  p99 {s['p99']:.2e}  max {s['max']:.2e}  frac>1e-3 {s['frac']:.4f}"
