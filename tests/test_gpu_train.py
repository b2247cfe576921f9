"""GPU parity of the training path: render_rays under autograd (csrc/train.cu's fp32 layer-by-layer
kernels and hand-written backward) against the oracle's torch-CPU autograd on the same seeded inputs and RNG draws.

Tolerances.  Forward outputs: same distribution criteria as the eval-mode parity tests.  Gradients: per parameter tensor,
cosine similarity with the oracle gradient and relative norm error; on the smooth test field (sigma head x5) cos >= 0.9998
and |norm ratio - 1| <= 3e-3 (24 rays; 0.997 / 1.5e-2 for the 12-ray variants); on the adversarial scene field (sigma head x40, where the reference's own outputs move by
> 1e-3 under one-ulp input changes) cos >= 0.98 on the whole flattened gradient."""
import pytest
import torch

from util import err_stats, fmt_stats

pytestmark = pytest.mark.gpu


def _models(sds, device="cuda"):
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    models = {}
    for name, sd in sds.items():
        has_n = "normal_net.0.weight" in sd
        has_m = "is_mirror_net.0.weight" in sd
        m = MirrorNeRF(predict_normal=has_n, predict_mirror_mask=has_m)
        m.load_state_dict(sd)
        models[name] = m.to(device).train()
    return models, {"xyz": Embedding(10), "dir": Embedding(4)}


def _smooth_sds(predict_normal=True, predict_mirror_mask=True):
    from mirror_nerf_b200.synthetic import make_state_dict
    return {"coarse": make_state_dict(21, 5.0, None, predict_normal, predict_mirror_mask),
            "fine": make_state_dict(22, 5.0, None, predict_normal, predict_mirror_mask)}


def _rng(n, Sc=64, Ni=128, seed=5):
    g = torch.Generator().manual_seed(seed)
    return {"perturb_u": torch.rand(n, Sc, generator=g), "noise_coarse": torch.randn(n, Sc, generator=g),
            "u_pdf": torch.rand(n, Ni, generator=g), "noise_fine": torch.randn(n, Sc + Ni, generator=g)}


def _loss(r, rays_d, seed=0):
    """A scalar that touches every differentiable output with fixed pseudo-random cotangents (stand-in for
    R/losses.py:201-259: colour, mirror-mask BCE, normal consistency, normal regularisation, plane consistency)."""
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for k in sorted(r):
        if k.startswith("z_vals"):
            continue
        v = r[k]
        c = torch.randn(v.shape, generator=g).to(v.device)
        total = total + (c * v).sum() + 0.5 * (v * v).sum()
    for typ in ("coarse", "fine"):
        if f"pred_normal_{typ}" in r:  # NormalRegLoss (losses.py:122-171)
            total = total + ((torch.relu(r[f"pred_normal_{typ}"] * rays_d.unsqueeze(1))).sum(-1) * r[f"weights_{typ}"]).mean()
    return total


def _grad_compare(models, params_cpu, cos_min, norm_tol, whole_cos_min=None):
    worst = (1.0, None)
    flat_a, flat_b = [], []
    for tag in params_cpu:
        named = dict(models[tag].named_parameters())
        for k, p in params_cpu[tag].items():
            if p.grad is None:
                assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, (tag, k)
                continue
            a = named[k].grad.detach().double().cpu().flatten()
            b = p.grad.detach().double().flatten()
            flat_a.append(a); flat_b.append(b)
            na, nb = float(a.norm()), float(b.norm())
            if nb < 1e-12:
                assert na < 1e-6, (tag, k, na)
                continue
            cos = float((a * b).sum() / (na * nb + 1e-300))
            if cos < worst[0]:
                worst = (cos, f"{tag}.{k}")
            if cos_min is not None:
                assert cos >= cos_min, (tag, k, cos, na, nb)
                assert abs(na / nb - 1.0) <= norm_tol, (tag, k, na, nb)
    a, b = torch.cat(flat_a), torch.cat(flat_b)
    whole = float((a * b).sum() / (a.norm() * b.norm()))
    print(f"gradient parity: worst per-tensor cos {worst[0]:.6f} ({worst[1]}), whole-vector cos {whole:.6f}, "
          f"norm ratio {float(a.norm() / b.norm()):.6f}")
    if whole_cos_min is not None:
        assert whole >= whole_cos_min, whole
    return whole


def _run_both(sds, n, kw_ours, args, seed=3, loss_seed=0):
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import mirror_nerf_oracle as O
    rays = random_rays(n, seed=seed)
    rng = _rng(n)
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    want = O.render_rays(params, rays, *args, rng=rng, **kw_ours)
    _loss(want, rays[:, 3:6], loss_seed).backward()
    models, emb = _models(sds)
    got = render_rays(models, emb, rays.cuda(), *args, rng=rng, **kw_ours)
    _loss(got, rays[:, 3:6].cuda(), loss_seed).backward()
    return got, want, models, params


def test_train_forward_and_gradients_smooth_field():
    """train.py:129-145 semantics: test_time=False, perturb=1, noise_std=1, compute_normal=True, both heads."""
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    got, want, models, params = _run_both(_smooth_sds(), 24, dict(test_time=False, compute_normal=True), args)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    assert torch.equal(got["z_vals_coarse"].cpu(), want["z_vals_coarse"])
    for k in sorted(got):
        assert tuple(got[k].shape) == tuple(want[k].shape), k
        assert got[k].requires_grad == want[k].requires_grad, k
        s = err_stats(got[k].detach().cpu(), want[k].detach())
        # analytic normals flip where a point sits on a ReLU kink: same allowance as the eval-path train-mode test
        # measured x 2 (profiles/r02_*_parity_stats.json: medians <= 3.0e-5; fraction beyond 1e-3: 5.6 % on the composited analytic
        # normals of this 24-ray batch, 1.2 % elsewhere)
        assert s["median"] <= 6e-5 and s["frac"] <= (0.11 if "normal" in k else 0.03), fmt_stats(k, s)
    _grad_compare(models, params, cos_min=0.9998, norm_tol=3e-3)  # measured 0.999958 / 1.3e-4, identical run to run


def test_train_gradients_adversarial_scene():
    from mirror_nerf_b200.synthetic import scene_state_dicts
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    got, want, models, params = _run_both(scene_state_dicts(), 16, dict(test_time=False, compute_normal=True), args)
    assert set(got) == set(want)
    _grad_compare(models, params, cos_min=None, norm_tol=None, whole_cos_min=0.98)


@pytest.mark.parametrize("variant", ["no_normal_grad", "white_back_disp", "coarse_only", "one_field", "no_heads",
                                     "detach_mask", "detach_normal", "detach_outside_mirror", "odd_sizes"])
def test_train_gradient_variants(variant):
    sds = _smooth_sds()
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    kw = dict(test_time=False, compute_normal=True)
    n = 12
    if variant == "no_normal_grad":
        kw["compute_normal"] = False
    elif variant == "white_back_disp":
        args = (64, True, 0.5, 0.0, 128, 32768, True)
    elif variant == "coarse_only":
        args = (64, False, 1.0, 1.0, 0, 32768, False)
        sds = {"coarse": sds["coarse"]}
    elif variant == "one_field":
        sds = {"coarse": sds["coarse"]}
        kw.update(only_one_field=True, current_epoch=5)
    elif variant == "no_heads":
        sds = _smooth_sds(False, False)
    elif variant == "detach_mask":
        kw["detach_density_for_mask_loss"] = True
    elif variant == "detach_normal":
        kw["detach_density_for_normal_loss"] = True
    elif variant == "odd_sizes":  # 7 rays x (40 + 25) samples: point counts that are not multiples of 16 / 128
        args = (40, False, 1.0, 1.0, 25, 32768, False)
        n = 7
    elif variant == "detach_outside_mirror":
        g = torch.Generator().manual_seed(9)
        kw.update(detach_density_outside_mirror_for_mask_loss=True,
                  mirror_mask=(torch.rand(n, generator=g) > 0.5).float())
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import mirror_nerf_oracle as O
    rays = random_rays(n, seed=4)
    rng = _rng(n, args[0], args[4] if args[4] else 1)
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    want = O.render_rays(params, rays, *args, rng=rng, **kw)
    _loss(want, rays[:, 3:6], 1).backward()
    models, emb = _models(sds)
    kw_gpu = dict(kw)
    if "mirror_mask" in kw_gpu:
        kw_gpu["mirror_mask"] = kw_gpu["mirror_mask"].cuda()
    got = render_rays(models, emb, rays.cuda(), *args, rng=rng, **kw_gpu)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    _loss(got, rays[:, 3:6].cuda(), 1).backward()
    # 12 rays only: one sample sitting on a ReLU kink (CPU and GPU round differently) moves the first layer's
    # second-order gradient; observed worst per-tensor cosine 0.9989 (xyz_encoding_1.0.weight) and norm ratio 1.0055
    _grad_compare(models, params, cos_min=0.997, norm_tol=1.5e-2)


def test_train_batch_split_and_accumulation(monkeypatch):
    """More rays than one autograd node takes: the groups' gradients accumulate to the single-node result."""
    from mirror_nerf_b200 import autograd as AG
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    sds = _smooth_sds()
    rays = random_rays(40, seed=8).cuda()
    rng = _rng(40)
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    grads = []
    for cap in (4096, 16):
        monkeypatch.setattr(AG, "MAX_TRAIN_RAYS", cap)
        models, emb = _models(sds)
        r = render_rays(models, emb, rays, *args, rng=rng, test_time=False, compute_normal=True)
        _loss(r, rays[:, 3:6], 2).backward()
        grads.append(torch.cat([p.grad.flatten() for m in models.values() for p in m.parameters()]).double())
    cos = float((grads[0] * grads[1]).sum() / (grads[0].norm() * grads[1].norm()))
    assert cos > 0.999999 and abs(float(grads[0].norm() / grads[1].norm()) - 1) < 1e-4, cos


def test_no_grad_path_unchanged_and_frozen_params():
    """Parameters that do not require grad (or torch.no_grad) keep using the tcgen05 inference path."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    models, emb = _models(_smooth_sds())
    rays = random_rays(8, seed=2).cuda()
    with torch.no_grad():
        r = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
    assert not r["rgb_fine"].requires_grad
    for m in models.values():
        for p in m.parameters():
            p.requires_grad_(False)
    r = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
    assert not r["rgb_fine"].requires_grad


def test_adam_kernel_matches_torch_adam():
    """mnrf_adam_step on flat buffers == torch.optim.Adam (R/utils/__init__.py:47-58: lr, eps=1e-8, L2 weight decay)."""
    from mirror_nerf_b200.mirror_nerf import MirrorNeRF
    from mirror_nerf_b200.parallel import FlatDataParallel
    from mirror_nerf_b200.synthetic import make_state_dict
    for wd in (0.0, 1e-2):
        ours = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        ours.load_state_dict(make_state_dict(3))
        ours = ours.cuda()
        ref = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        ref.load_state_dict(make_state_dict(3))
        ref = ref.cuda()
        ddp = FlatDataParallel({"coarse": ours}, lr=5e-4, weight_decay=wd)
        opt = torch.optim.Adam(ref.parameters(), lr=5e-4, eps=1e-8, weight_decay=wd)
        g = torch.Generator().manual_seed(0)
        for it in range(4):
            ddp.zero_grad()
            opt.zero_grad()
            for p, q in zip(ours.parameters(), ref.parameters()):
                gr = torch.randn(p.shape, generator=g).cuda() * (10.0 ** (it - 2))
                p.grad.add_(gr)
                q.grad = gr.clone()
            ddp.step()
            opt.step()
        for (k, p), q in zip(ours.named_parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), (wd, k, float((p - q).abs().max()))


def test_training_loop_reduces_loss_and_repacks_weights():
    """A few optimizer steps on one batch: the loss goes down, and the inference path sees the updated weights
    (FlatDataParallel.step invalidates the packed tensor-core weights)."""
    from mirror_nerf_b200.parallel import FlatDataParallel
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    models, emb = _models(_smooth_sds())
    ddp = FlatDataParallel(models, lr=1e-3)
    rays = random_rays(256, seed=12).cuda()
    g = torch.Generator().manual_seed(1)
    target = torch.rand(256, 3, generator=g).cuda()
    with torch.no_grad():
        before = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)["rgb_fine"].clone()
    losses = []
    for it in range(8):
        ddp.zero_grad()
        r = render_rays(models, emb, rays, 64, False, 1.0, 0.0, 128, 32768, False, test_time=False, compute_normal=True)
        loss = ((r["rgb_fine"] - target) ** 2).mean() + ((r["rgb_coarse"] - target) ** 2).mean() + 1e-4 * r["normal_dif_fine"].mean()
        loss.backward()
        ddp.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.9 * losses[0], losses
    with torch.no_grad():
        after = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)["rgb_fine"]
    assert float((after - before).abs().max()) > 1e-3


@pytest.mark.parametrize("compute_normal", [True, False])
def test_train_ray_gradients(compute_normal):
    """Rays that require grad (secondary rays of the training recursion, train.py:194-243) receive dL/d[o, d]: through the
    sample positions (PE Jacobian, incl. the second-order term of the analytic normals), the embedded direction and
    x_surface.  near / far get zero."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    from oracle import mirror_nerf_oracle as O
    sds = _smooth_sds()
    n = 16
    rays = random_rays(n, seed=6)
    rng = _rng(n)
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    kw = dict(test_time=False, compute_normal=compute_normal)
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    rc = rays.clone().requires_grad_(True)
    _loss(O.render_rays(params, rc, *args, rng=rng, **kw), rays[:, 3:6], 3).backward()
    models, emb = _models(sds)
    rg = rays.cuda().requires_grad_(True)
    _loss(render_rays(models, emb, rg, *args, rng=rng, **kw), rays[:, 3:6].cuda(), 3).backward()
    a, b = rg.grad.double().cpu(), rc.grad.double()
    assert float(a[:, 6:].abs().max()) == 0.0
    for name, sl in (("origin", slice(0, 3)), ("direction", slice(3, 6))):
        x, y = a[:, sl].flatten(), b[:, sl].flatten()
        cos = float((x * y).sum() / (x.norm() * y.norm()))
        print(f"ray gradient ({name}, compute_normal={compute_normal}): cos {cos:.6f}, norm ratio {float(x.norm() / y.norm()):.5f}")
        assert cos > 0.998 and abs(float(x.norm() / y.norm()) - 1) < 1.5e-2, (name, cos)  # measured 0.99924 / 0.4 %
    _grad_compare(models, params, cos_min=0.999, norm_tol=5e-3)
    # frozen parameters, rays only
    for m in models.values():
        for p in m.parameters():
            p.requires_grad_(False)
    rg2 = rays.cuda().requires_grad_(True)
    _loss(render_rays(models, emb, rg2, *args, rng=rng, **kw), rays[:, 3:6].cuda(), 3).backward()
    assert torch.allclose(rg2.grad, rg.grad, rtol=1e-3, atol=1e-5 * float(rg.grad.abs().max()))


def _train_recursion(render_fn, rays, level=0, max_level=1):
    """The reference's training recursion around render_rays (R/train.py:129-296, only_trace_rays_in_mirrors=True,
    predicted mask, nothing detached) in plain torch ops -- device agnostic, used for the oracle and for our render_rays."""
    r = render_fn(rays)
    mask = r["mirror_mask_fine"].detach().clone()
    mask[mask > 0.5] = 1
    mask[mask < 0.5] = 0
    if level >= max_level or not bool(mask.bool().any()):
        return r
    n = r["surface_normal_fine"]
    n = n / torch.sqrt(torch.clamp((n ** 2).sum(-1, keepdim=True), min=torch.finfo(torch.float32).eps))
    d = -rays[:, 3:6]
    w = d / torch.sqrt(torch.clamp((d ** 2).sum(-1, keepdim=True), min=torch.finfo(torch.float32).eps))
    cos = (w * n).sum(-1, keepdim=True)
    refl = 2 * cos * n - w
    sec = torch.cat([r["x_surface_fine"], refl, torch.full_like(rays[:, 6:7], 0.1), rays[:, 7:8]], -1)
    sec = sec[mask.bool()]
    child = _train_recursion(render_fn, sec, level + 1, max_level)
    for typ in ("coarse", "fine"):
        base = r[f"rgb_{typ}"]
        part = base.clone().detach()
        part[mask.bool()] = child[f"rgb_{typ}"]
        m3 = mask.float().unsqueeze(-1)
        r[f"rgb_{typ}_direct"] = base
        r[f"rgb_{typ}"] = m3 * part + (1 - m3) * base
    return r


def test_training_recursion_gradients_through_secondary_rays():
    """One bounce with train semantics: gradients reach the parameters through the child level's outputs AND through the
    secondary rays' geometry (x_surface, surface normal) -- the reference's callers need nothing but render_rays."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import make_state_dict, random_rays
    from oracle import mirror_nerf_oracle as O
    # smooth density, sharp mirror head so that a good part of the rays bounces
    sds = {"coarse": make_state_dict(21, 5.0, None, True, True, mirror_scale=100.0),
           "fine": make_state_dict(22, 5.0, None, True, True, mirror_scale=100.0)}
    n = 32
    rays = random_rays(n, seed=13)
    args = (64, False, 0.0, 0.0, 128, 32768, False)
    kw = dict(test_time=False, compute_normal=False)
    g = torch.Generator().manual_seed(4)
    target = torch.rand(n, 3, generator=g)
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    want = _train_recursion(lambda r: O.render_rays(params, r, *args, **kw), rays)
    assert "rgb_fine_direct" in want, "test scene must bounce"
    (((want["rgb_fine"] - target) ** 2).mean() + ((want["rgb_coarse"] - target) ** 2).mean()).backward()
    models, emb = _models(sds)
    got = _train_recursion(lambda r: render_rays(models, emb, r, *args, **kw), rays.cuda())
    (((got["rgb_fine"] - target.cuda()) ** 2).mean() + ((got["rgb_coarse"] - target.cuda()) ** 2).mean()).backward()
    s = err_stats(got["rgb_fine"].detach().cpu(), want["rgb_fine"].detach())
    assert s["median"] <= 1e-4 and s["frac"] <= 0.1, fmt_stats("rgb_fine (1 bounce)", s)
    _grad_compare(models, params, cos_min=None, norm_tol=None, whole_cos_min=0.995)


def test_train_full_size_engines_agree_and_gradients_are_linear():
    """BASELINE config-5 size (4096 rays x (64+128) samples, analytic normals), size-independent properties:
    (i) the tcgen05 tf32x3 GEMM engine and the fp32 CUDA-core twin produce the same gradients (cosine >= 0.9999 per large tensor);
    (ii) backward is linear in the cotangent: doubling the loss doubles every gradient (relative 1e-3; atomics reorder sums)."""
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    lib = _lib.load()
    sds = _smooth_sds()
    n = 4096
    rays = random_rays(n, seed=31).cuda()
    g = torch.Generator().manual_seed(2)
    rng = {k: v.cuda() for k, v in _rng(n).items()}
    target = torch.rand(n, 3, generator=g).cuda()
    args = (64, False, 1.0, 1.0, 128, 32768, False)

    def grads(engine, scale):
        _lib.check(lib.mnrf_train_set_gemm(engine))
        models, emb = _models(sds)
        r = render_rays(models, emb, rays, *args, rng=rng, test_time=False, compute_normal=True)
        loss = sum(((r[f"rgb_{t}"] - target) ** 2).mean() + 1e-2 * r[f"normal_dif_{t}"].mean() + r[f"mirror_mask_{t}"].mean()
                   + 1e-2 * (r[f"normal_{t}"] * r[f"weights_{t}"].unsqueeze(-1)).sum(1).pow(2).mean() for t in ("coarse", "fine"))
        (scale * loss).backward()
        return {f"{t}.{k}": p.grad.double() for t, m in models.items() for k, p in m.named_parameters()}, float(loss.detach())

    try:
        g_tc, l_tc = grads(1, 1.0)
        g_tc2, _ = grads(1, 2.0)
        g_simt, l_simt = grads(0, 1.0)
    finally:
        lib.mnrf_train_set_gemm(1)
    assert abs(l_tc - l_simt) <= 1e-4 * abs(l_simt), (l_tc, l_simt)
    worst = 1.0
    for k in g_tc:
        a, b, a2 = g_tc[k].flatten(), g_simt[k].flatten(), g_tc2[k].flatten()
        if float(b.norm()) < 1e-12:
            continue
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        worst = min(worst, cos)
        if a.numel() >= 256:
            assert cos >= 0.9999, (k, cos)
        assert float((a2 - 2 * a).norm()) <= 1e-3 * float((2 * a).norm()) + 1e-12, k
    print(f"full-size engine agreement: worst per-tensor cosine {worst:.6f}")


def test_train_single_pass_tf32_speed_mode():
    """Engine 2 (one tf32 pass, opt-in speed mode; 39.0 vs 45.4 ms per 4096-ray step): gradients stay in the neighbourhood of the
    fp32-grade default (whole-vector cosine >= 0.99; measured 0.9958 on 64 rays) but are measurably different -- this is why
    the 3-pass split is the default and the only parity mode."""
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    lib = _lib.load()
    sds = _smooth_sds()
    rays = random_rays(64, seed=17).cuda()
    rng = _rng(64)
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    out = []
    try:
        for engine in (1, 2):
            _lib.check(lib.mnrf_train_set_gemm(engine))
            models, emb = _models(sds)
            r = render_rays(models, emb, rays, *args, rng=rng, test_time=False, compute_normal=True)
            _loss(r, rays[:, 3:6], 4).backward()
            out.append(torch.cat([p.grad.flatten() for m in models.values() for p in m.parameters()]).double())
    finally:
        lib.mnrf_train_set_gemm(1)
    cos = float((out[0] * out[1]).sum() / (out[0].norm() * out[1].norm()))
    print(f"tf32x1 vs tf32x3 gradient cosine {cos:.6f}")
    assert 0.99 <= cos < 1.0 - 1e-9  # close, but measurably not the same arithmetic


def test_recursive_eval_driver_refuses_to_run_under_autograd():
    from mirror_nerf_b200.synthetic import random_rays
    from mirror_nerf_b200.trace import render_rays_recursive
    models, emb = _models(_smooth_sds())
    rays = random_rays(8, seed=1).cuda()
    with pytest.raises(NotImplementedError, match="inference"):
        render_rays_recursive(models, emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
    with torch.no_grad():
        r = render_rays_recursive(models, emb, rays, 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
    assert "rgb_fine" in r


def test_train_gradients_on_the_fitted_room_field():
    """Gradient parity on the scene-like field (tests/golden/room_field.npz) with rays of the analytic room and the training
    loss of tools/train_room.py: per-tensor cosine >= 0.9999, norm within 2e-2."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.room_scene import random_room_rays, trace_room
    from oracle import mirror_nerf_oracle as O
    from util import room_state_dicts
    sds = room_state_dicts()
    n = 24
    rays = random_room_rays(n, torch.Generator().manual_seed(3))
    gt, mask_gt, _ = trace_room(rays)
    rng = _rng(n)
    args = (64, False, 1.0, 0.0, 128, 32768, False)
    kw = dict(test_time=False, compute_normal=True)

    def loss_of(r, dev):
        tot = 0.0
        for typ in ("coarse", "fine"):
            tot = tot + ((r[f"rgb_{typ}"] - gt.to(dev)) ** 2).mean()
            tot = tot + 0.1 * torch.nn.functional.binary_cross_entropy(r[f"mirror_mask_{typ}"].clamp(1e-7, 1 - 1e-7), mask_gt.to(dev))
            tot = tot + 1e-2 * r[f"normal_dif_{typ}"].mean()
        return tot
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    lw = loss_of(O.render_rays(params, rays, *args, rng=rng, **kw), "cpu")
    lw.backward()
    models, emb = _models(sds)
    lg = loss_of(render_rays(models, emb, rays.cuda(), *args, rng=rng, **kw), "cuda")
    lg.backward()
    assert abs(float(lg) - float(lw)) <= 1e-4 * abs(float(lw)), (float(lg), float(lw))
    # sharp fitted surfaces: the norms of the small normal-loss gradients move by up to 0.6 % (normal_net.0.weight), directions agree
    _grad_compare(models, params, cos_min=0.9999, norm_tol=2e-2)


def test_functional_training_on_the_room_scene():
    """The whole training stack on the analytic mirror-room scene (mirror_nerf_b200/room_trainer.py): 250 optimizer steps of
    2048 rays with the one-bounce train-time recursion raise the held-out PSNR by more than 4 dB and teach the mirror mask
    (gradients are accumulated with atomics, so runs differ: observed gains after 200 steps 4.5 .. 5.6 dB)."""
    from mirror_nerf_b200.room_trainer import RoomTrainer
    tr = RoomTrainer(rays_per_step=2048, seed=0)
    p0, _ = tr.psnr(res=96)
    for _ in range(250):
        loss = tr.step()
    assert torch.isfinite(loss)
    p1, mirror_frac = tr.psnr(res=96)
    print(f"room scene, 250 steps: PSNR {p0:.2f} -> {p1:.2f} dB, predicted mirror fraction {mirror_frac:.3f}")
    assert p1 > p0 + 4.0, (p0, p1)
    assert 0.2 < mirror_frac < 0.65  # ground truth of this view: 0.42
