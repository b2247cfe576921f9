"""Generate golden vectors by running the UNMODIFIED reference (needs /root/reference; build container only).

    python tests/golden/make_golden.py

Imports R/models/mirror_nerf.py + R/models/rendering.py (with a stub for the missing, unused
``torch_optimizer`` import of R/utils/__init__.py:7), runs them on the seeded inputs of
``mirror_nerf_b200.synthetic`` and writes tests/golden/*.npz.  The committed .npz files travel to the
GPU box; this script and /root/reference do not need to.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MNRF_REFERENCE", "/root/reference")


def import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("torch_optimizer", types.ModuleType("torch_optimizer"))
    from models.mirror_nerf import Embedding, MirrorNeRF  # noqa
    from models.rendering import render_rays, sample_pdf  # noqa
    return Embedding, MirrorNeRF, render_rays, sample_pdf


def npify(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def generate():
    """Run the reference on the seeded inputs; returns {fixture name: {key: ndarray}} (nothing is written)."""
    from mirror_nerf_b200.synthetic import random_rays, scene_state_dicts
    torch.set_num_threads(8)
    Embedding, MirrorNeRF, render_rays, sample_pdf = import_reference()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}

    def model(sd):
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        return m.eval()

    sds = scene_state_dicts()
    coarse, fine = model(sds["coarse"]), model(sds["fine"])
    models = {"coarse": coarse, "fine": fine}

    # ---- 1. embedding + field forward on flat points ------------------------------------------
    g = np.random.Generator(np.random.PCG64(7))
    xyz = torch.from_numpy(g.uniform(-3.0, 3.0, size=(192, 3)).astype(np.float32))
    xyz[0] = 0.0
    xyz[1] = torch.tensor([8.0, -8.0, 7.5])
    d = torch.from_numpy(g.standard_normal(size=(192, 3)).astype(np.float32))
    d = d / d.norm(dim=-1, keepdim=True)
    out = {"xyz": xyz, "dir": d, "pe_xyz": emb["xyz"](xyz), "pe_dir": emb["dir"](d)}
    x_full = torch.cat([xyz, out["pe_dir"]], 1)
    with torch.no_grad():
        o = fine(x_full.clone(), compute_normal=False, sigma_only=False, embedding_xyz=emb["xyz"])
    for k in ("sigma", "geo_feat", "pred_normal", "rgb", "is_mirror"):
        out["full_" + k] = o[k]
    with torch.no_grad():
        o = fine(xyz.clone(), compute_normal=False, sigma_only=True, embedding_xyz=emb["xyz"])
    for k in ("sigma", "pred_normal"):
        out["sigonly_" + k] = o[k]
    o = fine(x_full.clone(), compute_normal=True, sigma_only=False, embedding_xyz=emb["xyz"])
    for k in ("sigma", "normal", "pred_normal", "rgb", "is_mirror"):
        out["grad_" + k] = o[k]
    files = {"field": npify(out)}

    # ---- 2. sample_pdf ---------------------------------------------------------------------------
    g = np.random.Generator(np.random.PCG64(11))
    n = 64
    z = np.sort(g.uniform(0.05, 8.0, size=(n, 64)).astype(np.float32), axis=1)
    bins = torch.from_numpy(0.5 * (z[:, 1:] + z[:, :-1]))
    w = g.uniform(0, 1, size=(n, 62)).astype(np.float32) ** 8
    w[0] = 0.0                      # all-zero row -> uniform pdf
    w[1] = 0.0; w[1, 17] = 1.0      # single spike
    w[2] = 0.0; w[2, 0] = 0.7; w[2, 61] = 0.3
    w[3] = 1e-7
    w = torch.from_numpy(w)
    det = sample_pdf(bins, w, 128, det=True)
    torch.manual_seed(123)
    rnd = sample_pdf(bins, w, 128, det=False)
    torch.manual_seed(123)
    u = torch.rand(n, 128)
    # recover the indices exactly as the reference computes them (R/models/rendering.py:19-35)
    ww = w + 1e-5
    pdf = ww / ww.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(n, 1), torch.cumsum(pdf, -1)], -1)
    u_det = torch.linspace(0, 1, 128).expand(n, 128).contiguous()
    files["sample_pdf"] = npify({
        "bins": bins, "weights": w, "det": det, "rnd": rnd, "u": u, "cdf": cdf,
        "inds_det": torch.searchsorted(cdf, u_det, right=True),
        "inds_rnd": torch.searchsorted(cdf, u, right=True)})

    # ---- 3. render_rays, eval mode 64+128 (BASELINE config 2 per-level call) ---------------------
    rays = random_rays(64, seed=1)
    with torch.no_grad():
        r = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True,
                        compute_normal=False)
    files["render_eval"] = dict(rays=rays.numpy(), **npify(r))

    # ---- 4. variants -------------------------------------------------------------------------------
    rays_s = random_rays(12, seed=2)
    var = {"rays": rays_s}

    def put(tag, res):
        for k, v in res.items():
            var[f"{tag}/{k}"] = v

    with torch.no_grad():
        put("white_back", render_rays(models, emb, rays_s, 64, False, 0, 0, 128, 32768, True, test_time=True,
                                      compute_normal=False))
        put("use_disp", render_rays(models, emb, rays_s, 64, True, 0, 0, 128, 32768, False, test_time=True,
                                    compute_normal=False))
        put("coarse_only", render_rays({"coarse": coarse}, emb, rays_s, 64, False, 0, 0, 0, 32768, False,
                                       test_time=True, compute_normal=False))
        put("one_field", render_rays({"coarse": coarse}, emb, rays_s, 64, False, 0, 0, 128, 32768, False,
                                     test_time=True, compute_normal=False, only_one_field=True,
                                     current_epoch=3))
        put("train_nonormal", render_rays(models, emb, rays_s, 64, False, 0, 0, 128, 32768, False,
                                          test_time=False, compute_normal=False))
        put("s32_i16", render_rays(models, emb, rays_s, 32, False, 0, 0, 16, 1000, False, test_time=True,
                                   compute_normal=False))
    files["render_variants"] = npify(var)

    # ---- 5. train mode: grad normals, perturb + noise with a replayable RNG stream, and gradients --
    rays_t = random_rays(8, seed=3)
    torch.manual_seed(5)
    rng = {"perturb_u": torch.rand(8, 64), "noise_coarse": torch.randn(8, 64), "u_pdf": torch.rand(8, 128),
           "noise_fine": torch.randn(8, 192)}
    for m in (coarse, fine):
        m.zero_grad()
    torch.manual_seed(5)
    r = render_rays(models, emb, rays_t, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False,
                    compute_normal=True)
    tr = {"rays": rays_t}
    tr.update({f"rng/{k}": v for k, v in rng.items()})
    tr.update({f"out/{k}": v for k, v in r.items()})
    # a loss touching every differentiable output (stand-in for R/losses.py:201-259)
    loss = sum((r[f"rgb_{t}"] ** 2).sum() + r[f"mirror_mask_{t}"].sum() + 0.1 * r[f"normal_dif_{t}"].sum()
               + 0.01 * (r[f"depth_{t}"]).sum() for t in ("coarse", "fine"))
    loss.backward()
    tr["loss"] = loss.detach()
    for tag, m in (("coarse", coarse), ("fine", fine)):
        for k, p in m.named_parameters():
            # big matrices: every 13th element + the L2 norm (keeps the fixture small)
            tr[f"grad/{tag}/{k}"] = p.grad.flatten()[::13] if p.grad.numel() > 4096 else p.grad
            tr[f"gradnorm/{tag}/{k}"] = p.grad.norm()
    files["render_train"] = npify(tr)
    return files


def main():
    for name, d in generate().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
