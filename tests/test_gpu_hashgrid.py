"""GPU parity of the hash-grid field (BASELINE config 3) against oracle/hashgrid_oracle.py.
The oracle restates tinycudann's published HashGrid algorithm (PARITY UNPINNED: the reference's encoder is an un-vendored,
unpinned third-party CUDA extension -- see the oracle's header); what these tests pin is kernel == restatement.
Tolerances: table indices are integer work (any mismatch shows up as O(1) feature errors); per-point outputs median <= 1e-5,
p99 <= 1e-3 relative (fp32 on both sides, different summation order); rendered outputs like the MLP-field eval tests."""
import pytest
import torch

from util import err_stats, fmt_stats

pytestmark = pytest.mark.gpu


def _module(sd, bound=1.0):
    from mirror_nerf_b200.mirror_nerf_tcnn import MirrorNeRFTcnn
    m = MirrorNeRFTcnn(bound=bound, predict_normal="normal_net.0.weight" in sd,
                       predict_mirror_mask="is_mirror_net.0.weight" in sd)
    m.load_state_dict(sd)
    return m.cuda().eval()


def _close(a, b, name, median=1e-5, p99=1e-3):
    s = err_stats(a.cpu(), b)
    assert s["median"] <= median and s["p99"] <= p99, fmt_stats(name, s)


@pytest.mark.parametrize("bound", [1.0, 2.0])
def test_hash_field_points_vs_oracle(bound):
    from oracle import hashgrid_oracle as H
    sd = H.make_state_dict(3, bound=bound)
    m = _module(sd, bound)
    g = torch.Generator().manual_seed(1)
    n = 4096
    xyz = (torch.rand(n, 3, generator=g) * 2 - 1) * bound * 1.05  # a few points outside the box (negative grid coordinates)
    xyz[0] = 0.0
    xyz[1] = bound
    xyz[2] = -bound
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    x = torch.cat([xyz, d], 1)
    want = H.field_forward(sd, x, bound=bound)
    with torch.no_grad():
        got = m(x.cuda(), compute_normal=False)
    assert set(got) == {"sigma", "pred_normal", "rgb", "is_mirror"}
    _close(got["sigma"], want["sigma"].flatten(), "sigma")
    _close(got["rgb"], want["rgb"], "rgb")
    _close(got["pred_normal"], want["pred_normal"], "pred_normal", median=2e-5, p99=5e-3)
    _close(got["is_mirror"], want["is_mirror"], "is_mirror")
    with torch.no_grad():
        so = m(xyz.cuda(), compute_normal=False, sigma_only=True)
    assert set(so) == {"sigma", "pred_normal"}
    _close(so["sigma"], want["sigma"].flatten(), "sigma (sigma_only)")
    # analytic normals: kernel's explicit derivative of the trilinear weights vs autograd through the restated encoder
    want_n = H.field_forward(sd, x, bound=bound, compute_normal=True)["normal"]
    with torch.no_grad():
        got_n = m(x.cuda(), compute_normal=True)["normal"]
    cos = (got_n.cpu() * want_n).sum(-1)
    assert float(cos.median()) > 1 - 1e-6 and float((cos < 1 - 1e-3).float().mean()) <= 0.01, (float(cos.min()), float(cos.median()))


def test_hash_field_without_heads_and_errors():
    from oracle import hashgrid_oracle as H
    sd = H.make_state_dict(5, predict_normal=False, predict_mirror_mask=False)
    m = _module(sd)
    x = torch.rand(64, 6) * 2 - 1
    want = H.field_forward(sd, x)
    with torch.no_grad():
        got = m(x.cuda(), compute_normal=False)
    assert set(got) == {"sigma", "rgb"}
    _close(got["rgb"], want["rgb"], "rgb")
    with torch.no_grad():
        gn = m(x.cuda())  # compute_normal defaults to True like the reference
    assert set(gn) == {"sigma", "rgb", "normal"}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, compute_normal=False)


def test_render_rays_hash_field_vs_oracle():
    """render_rays with the nerf_tcnn model family and identity embeddings (R/train.py:69-100), eval mode 64+128."""
    from mirror_nerf_b200.mirror_nerf import Embedding
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import hashgrid_oracle as H
    from oracle import mirror_nerf_oracle as O
    sds = {"coarse": H.make_state_dict(7, sigma_scale=20.0), "fine": H.make_state_dict(8, sigma_scale=20.0)}
    models = {k: _module(v) for k, v in sds.items()}
    emb = {"xyz": Embedding(0), "dir": Embedding(0)}
    rays = random_rays(96, seed=5, near=0.05, far=2.0)
    want = O.render_rays(sds, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False,
                         n_freqs_xyz=0, n_freqs_dir=0)
    with torch.no_grad():
        got = render_rays(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False, test_time=True,
                          compute_normal=False)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    assert torch.equal(got["z_vals_coarse"].cpu(), want["z_vals_coarse"])
    for k in sorted(got):
        assert tuple(got[k].shape) == tuple(want[k].shape), k
        s = err_stats(got[k].cpu(), want[k])
        assert s["median"] <= 1e-4 and s["frac"] <= 0.05, fmt_stats(k, s)
    # compute_normal=True (the signature's default): analytic normals through the hash grid
    want_n = O.render_rays(sds, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=True,
                           n_freqs_xyz=0, n_freqs_dir=0)
    with torch.no_grad():
        got_n = render_rays(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False, test_time=True)
    assert set(got_n) == set(want_n), sorted(set(got_n) ^ set(want_n))
    for k in ("normal_fine", "surface_normal_grad_fine", "normal_dif_fine", "rgb_fine"):
        s = err_stats(got_n[k].cpu(), want_n[k])
        # analytic normals jump at cell faces of the fine levels and at ReLU kinks: same allowance as the MLP-field tests
        assert s["median"] <= 1e-4 and s["frac"] <= (0.13 if "normal" in k else 0.05), fmt_stats(k, s)
    with pytest.raises(NotImplementedError):
        render_rays(models, {"xyz": Embedding(10), "dir": Embedding(4)}, rays.cuda(), 64, False, 0, 0, 128, 32768, False,
                    test_time=True, compute_normal=False)


# ---- training: gradients of the hash-grid field through render_rays (csrc/train_hash.cu) ----------------------------------------
def _train_models(sds, bound=1.0):
    from mirror_nerf_b200.mirror_nerf import Embedding
    models = {k: _module(v, bound).train() for k, v in sds.items()}
    return models, {"xyz": Embedding(0), "dir": Embedding(0)}


def _hash_sds(predict_normal=True, predict_mirror_mask=True, sigma_scale=6.0):
    from oracle import hashgrid_oracle as H
    return {"coarse": H.make_state_dict(31, sigma_scale=sigma_scale, predict_normal=predict_normal,
                                        predict_mirror_mask=predict_mirror_mask),
            "fine": H.make_state_dict(32, sigma_scale=sigma_scale, predict_normal=predict_normal,
                                      predict_mirror_mask=predict_mirror_mask)}


HASH_TRAIN_VARIANTS = {
    "full": dict(kw=dict(compute_normal=True)),
    "no_analytic_normal": dict(kw=dict(compute_normal=False)),
    "detach_flags": dict(kw=dict(compute_normal=True, detach_density_for_mask_loss=True, detach_density_for_normal_loss=True)),
    "detach_outside_mirror": dict(kw=dict(compute_normal=True, detach_density_outside_mirror_for_mask_loss=True), gt_mask=True),
    "no_heads": dict(kw=dict(compute_normal=True), heads=False),
    "white_back_coarse_only": dict(kw=dict(compute_normal=True), args=(24, False, 1.0, 1.0, 0, 32768, True)),
}


@pytest.mark.parametrize("variant", list(HASH_TRAIN_VARIANTS))
def test_hash_train_gradients_vs_oracle(variant):
    """R/train.py:129-145 with --model_type nerf_tcnn: test_time=False, perturb=1, noise_std=1.  Forward outputs and the
    gradients of every parameter tensor (hash table included) against torch autograd over the oracle restatement.
    Forward: an independent oracle run.  Gradients: a second oracle run at the implementation's fine depths (oracle hook
    rng["z_fine"]; the reference detaches them, R/models/rendering.py:346-349) -- independently sampled depths differ in the last
    bits, which moves a few of the 2,368 samples across cell faces of the fine grid levels where the trilinear Jacobian jumps
    (~2e-4 crossings per sample, axis and level: per-tensor cosine 0.9995..0.99994 depending on the level table's last bits),
    a property of the encoding, not of either implementation."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import mirror_nerf_oracle as O
    from test_gpu_train import _grad_compare, _loss, _rng
    v = HASH_TRAIN_VARIANTS[variant]
    heads = v.get("heads", True)
    args = v.get("args", (24, False, 1.0, 1.0, 40, 32768, False))
    kw = dict(v["kw"], test_time=False)
    n = 37
    sds = _hash_sds(heads, heads)
    rays = random_rays(n, seed=9, near=0.05, far=2.0)
    rng = _rng(n, args[0], args[4])
    if v.get("gt_mask"):
        kw["mirror_mask"] = (torch.arange(n) % 2).float()
    params = {t: {k: x.clone().requires_grad_(True) for k, x in sd.items()} for t, sd in sds.items()}
    with torch.no_grad():
        want = O.render_rays(sds, rays, *args, rng=rng, n_freqs_xyz=0, n_freqs_dir=0, **kw)
    models, emb = _train_models(sds)
    kw_gpu = dict(kw)
    if "mirror_mask" in kw_gpu:
        kw_gpu["mirror_mask"] = kw_gpu["mirror_mask"].cuda()
    got = render_rays(models, emb, rays.cuda(), *args, rng=rng, **kw_gpu)
    _loss(got, rays[:, 3:6].cuda(), 1).backward()
    rng_g = dict(rng)
    if "z_vals_fine" in got:
        rng_g["z_fine"] = got["z_vals_fine"].detach().cpu()
    want_g = O.render_rays(params, rays, *args, rng=rng_g, n_freqs_xyz=0, n_freqs_dir=0, **kw)
    _loss(want_g, rays[:, 3:6], 1).backward()
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    for k in sorted(got):
        assert tuple(got[k].shape) == tuple(want[k].shape), k
        s = err_stats(got[k].detach().cpu(), want[k].detach())
        assert s["median"] <= 1e-4 and s["frac"] <= (0.13 if "normal" in k else 0.05), fmt_stats(k, s)
    if args[4] == 0:
        params = {"coarse": params["coarse"]}
    # measured at identical depths: every tensor and the whole vector 1.000000 (6 digits), norm ratio 1.000000
    _grad_compare(models, params, 0.99999, 1e-4, whole_cos_min=0.99999)


@pytest.mark.parametrize("compute_normal", [True, False])
def test_hash_train_ray_gradients(compute_normal):
    """Secondary rays are built from x_surface / normals without detaching (R/train.py:194-243): d L / d [o, d] through the hash
    encoding (first order, and the mixed second derivatives of the trilinear weights under the analytic normal), the SH
    direction encoding and x_surface."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import mirror_nerf_oracle as O
    from test_gpu_train import _loss, _rng
    n, args = 29, (24, False, 1.0, 1.0, 40, 32768, False)
    sds = _hash_sds()
    rays = random_rays(n, seed=10, near=0.05, far=2.0)
    rng = _rng(n, args[0], args[4])
    kw = dict(test_time=False, compute_normal=compute_normal)
    models, emb = _train_models(sds)
    for m in models.values():
        m.requires_grad_(False)
    rg = rays.cuda().requires_grad_(True)
    got = render_rays(models, emb, rg, *args, rng=rng, **kw)
    _loss(got, rays[:, 3:6].cuda(), 2).backward()
    # the oracle at the implementation's (detached) fine depths: see test_hash_train_gradients_vs_oracle
    rc = rays.clone().requires_grad_(True)
    want = O.render_rays(sds, rc, *args, rng=dict(rng, z_fine=got["z_vals_fine"].detach().cpu()), n_freqs_xyz=0, n_freqs_dir=0, **kw)
    _loss(want, rays[:, 3:6], 2).backward()
    a, b = rg.grad[:, :6].double().cpu().flatten(), rc.grad[:, :6].double().flatten()
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    print(f"hash-grid ray gradient: cos {cos:.6f}, norm ratio {float(a.norm() / b.norm()):.6f}")
    assert cos >= 0.99999 and abs(float(a.norm() / b.norm()) - 1) <= 1e-4, (cos, float(a.norm()), float(b.norm()))  # measured 1.000000 / 1.000000
    assert float(rg.grad[:, 6:].abs().max()) == 0.0


def test_hash_training_loop_reduces_loss_and_repacks_table():
    """A few Adam steps on both hash-grid fields: the loss falls and every step re-packs the changed table in place."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    models, emb = _train_models(_hash_sds(sigma_scale=3.0))
    params = [p for m in models.values() for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=1e-2)
    rays = random_rays(256, seed=12, near=0.05, far=2.0).cuda()
    target = torch.rand(256, 3, generator=torch.Generator().manual_seed(0)).cuda()
    handles, losses = set(), []
    for step in range(12):
        opt.zero_grad(set_to_none=True)
        r = render_rays(models, emb, rays, 32, False, 1.0, 0.0, 32, 32768, False, test_time=False, compute_normal=True)
        loss = ((r["rgb_fine"] - target) ** 2).mean() + ((r["rgb_coarse"] - target) ** 2).mean() + 1e-3 * r["normal_dif_fine"].mean()
        loss.backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in params)
        opt.step()
        losses.append(float(loss.detach()))
        handles.add(models["fine"].__dict__["_mnrf_packed"][0].handle.value)
    assert len(handles) == 1, "the packed field must be updated in place, not re-created"
    assert losses[-1] < 0.8 * losses[0], losses
