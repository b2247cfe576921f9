"""CPU oracle for the Mirror-NeRF render_rays hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (functional, dict-of-tensors, no nn.Module) of the
reference algorithm.  It exists so that tests, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have something to check the CUDA path
against on a box where ``/root/reference`` does not exist.  Nothing in ``mirror_nerf_b200/`` (the
product) imports it, and it must never be used as a fallback for a missing CUDA extension.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
(``/root/reference/models/{rendering,mirror_nerf}.py``) in the build container, runs it on seeded
inputs and stores inputs+outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this restatement against those vectors (bit-exact for indices, <=1e-6 for fp32 values; in the
container that produced them every tensor is bit-identical).

Reference lines followed (R/ = zju3dv/Mirror-NeRF @ fc0d7911):
  embed            R/models/mirror_nerf.py:6-38
  l2_normalize     R/utils/func.py:5-7
  trunk            R/models/mirror_nerf.py:189-197
  field_forward    R/models/mirror_nerf.py:101-187, 199-212 ; grad normal R/utils/func.py:10-25
  sample_pdf       R/models/rendering.py:7-51
  composite        R/models/rendering.py:175-264
  render_rays      R/models/rendering.py:54-369
  trace_level      R/train.py:129-348 (train semantics) / R/eval.py:132-725 (eval semantics)
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

FP32_EPS = 1.1920928955078125e-07  # torch.finfo(torch.float32).eps, R/utils/func.py:5


# --------------------------------------------------------------------------------------------
# parameter container (same key names / [out,in] layout as MirrorNeRF.state_dict(),
# R/models/mirror_nerf.py:60-99 and SURVEY.md section 5 "Checkpoint")
# --------------------------------------------------------------------------------------------
def param_shapes(D=8, W=256, in_xyz=63, in_dir=27, skips=(4,), predict_normal=True,
                 predict_mirror_mask=True):
    """Ordered {state_dict key: shape} for a MirrorNeRF field."""
    s = OrderedDict()
    for i in range(D):
        fan_in = in_xyz if i == 0 else (W + in_xyz if i in skips else W)
        s[f"xyz_encoding_{i + 1}.0.weight"] = (W, fan_in)
        s[f"xyz_encoding_{i + 1}.0.bias"] = (W,)
    s["xyz_encoding_final.weight"] = (W, W)
    s["xyz_encoding_final.bias"] = (W,)
    s["dir_encoding.0.weight"] = (W // 2, W + in_dir)
    s["dir_encoding.0.bias"] = (W // 2,)
    s["sigma.weight"] = (1, W)
    s["sigma.bias"] = (1,)
    s["rgb.0.weight"] = (3, W // 2)
    s["rgb.0.bias"] = (3,)
    if predict_normal:
        s["normal_net.0.weight"] = (W // 2, W)
        s["normal_net.0.bias"] = (W // 2,)
        s["normal_net.1.weight"] = (3, W // 2)
        s["normal_net.1.bias"] = (3,)
    if predict_mirror_mask:
        s["is_mirror_net.0.weight"] = (W // 2, W)
        s["is_mirror_net.0.bias"] = (W // 2,)
        s["is_mirror_net.2.weight"] = (1, W // 2)
        s["is_mirror_net.2.bias"] = (1,)
    return s


# --------------------------------------------------------------------------------------------
# field
# --------------------------------------------------------------------------------------------
def embed(x, n_freqs):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (R/models/mirror_nerf.py:33-38)."""
    if n_freqs == 0:
        return torch.cat([x], -1)
    bands = 2 ** torch.linspace(0, n_freqs - 1, n_freqs)  # :17 (logscale=True)
    parts = [x]
    for f in bands:
        parts.append(torch.sin(f * x))
        parts.append(torch.cos(f * x))
    return torch.cat(parts, -1)


def l2_normalize(x):
    """x / sqrt(max(sum x^2, eps)) -- the clamp is on the SQUARED norm (R/utils/func.py:5-7)."""
    return x / torch.sqrt(torch.maximum(torch.sum(x ** 2, dim=-1, keepdim=True),
                                        torch.as_tensor(FP32_EPS)))


def trunk(p, pe, D=8, skips=(4,)):
    """8 x (Linear+ReLU) with [input, h] skip concat before layer 5; raw sigma head."""
    h = pe
    for i in range(D):
        if i in skips:
            h = torch.cat([pe, h], -1)
        h = torch.relu(F.linear(h, p[f"xyz_encoding_{i + 1}.0.weight"], p[f"xyz_encoding_{i + 1}.0.bias"]))
    sigma = F.linear(h, p["sigma.weight"], p["sigma.bias"])
    return sigma, h


def _color(p, geo, dir_emb):
    f = F.linear(geo, p["xyz_encoding_final.weight"], p["xyz_encoding_final.bias"])
    h = torch.relu(F.linear(torch.cat([f, dir_emb], -1), p["dir_encoding.0.weight"], p["dir_encoding.0.bias"]))
    return torch.sigmoid(F.linear(h, p["rgb.0.weight"], p["rgb.0.bias"]))


def _pred_normal(p, geo):
    h = F.linear(geo, p["normal_net.0.weight"], p["normal_net.0.bias"])  # NO activation (:85-88)
    return F.linear(h, p["normal_net.1.weight"], p["normal_net.1.bias"])


def _is_mirror(p, geo):
    h = F.leaky_relu(F.linear(geo, p["is_mirror_net.0.weight"], p["is_mirror_net.0.bias"]), 0.01)
    return torch.sigmoid(F.linear(h, p["is_mirror_net.2.weight"], p["is_mirror_net.2.bias"]))


def field_forward(p, x, *, n_freqs_xyz=10, in_dir=27, compute_normal=True, sigma_only=False,
                  mirror_mask=None, detach_density_outside_mirror_for_mask_loss=False,
                  detach_density_for_mask_loss=False, detach_density_for_normal_loss=False):
    """MirrorNeRF.forward on a flat batch (R/models/mirror_nerf.py:101-187).

    x: (B, 3+in_dir) = [xyz | embedded dir]  or (B,3) when sigma_only.
    Returns the same dict keys: sigma (B,1), geo_feat, normal?, pred_normal?, rgb?, is_mirror?.
    """
    out = {}
    if not sigma_only:
        xyz, dir_emb = torch.split(x, [3, in_dir], dim=-1)
    else:
        xyz = x
    if compute_normal:
        xyz.requires_grad_(True)
        with torch.enable_grad():
            sigma, geo = trunk(p, embed(xyz, n_freqs_xyz))
        (g,) = torch.autograd.grad(sigma, xyz, torch.ones_like(sigma), create_graph=True,
                                   retain_graph=True, only_inputs=True)
        out["normal"] = l2_normalize(-g)
    else:
        sigma, geo = trunk(p, embed(xyz, n_freqs_xyz))
    out["sigma"] = sigma
    out["geo_feat"] = geo
    if "normal_net.0.weight" in p:
        out["pred_normal"] = l2_normalize(
            _pred_normal(p, geo.detach() if detach_density_for_normal_loss else geo))
    if not sigma_only:
        out["rgb"] = _color(p, geo, dir_emb)
        if "is_mirror_net.0.weight" in p:
            if detach_density_for_mask_loss:
                out["is_mirror"] = _is_mirror(p, geo.detach())
            elif (detach_density_outside_mirror_for_mask_loss and mirror_mask is not None
                  and not bool((mirror_mask < 0).any())):
                keep = mirror_mask.clone().bool()
                g2 = geo.clone()
                g2[~keep] = g2[~keep].detach()
                out["is_mirror"] = _is_mirror(p, g2)
            else:
                out["is_mirror"] = _is_mirror(p, geo)
    return out


def analytic_normal_explicit(p, xyz, n_freqs=10, D=8, skips=(4,)):
    """-d sigma / d xyz (un-normalised gradient, sign NOT flipped) by an explicit reverse chain.

    Restates what autograd does for R/models/mirror_nerf.py:136-146 (ReLU' masks, skip split, PE
    Jacobian); used to pin the CUDA kernel's hand-written chain against autograd.
    """
    pe = embed(xyz, n_freqs)
    hs, h = [], pe
    for i in range(D):
        if i in skips:
            h = torch.cat([pe, h], -1)
        h = torch.relu(F.linear(h, p[f"xyz_encoding_{i + 1}.0.weight"], p[f"xyz_encoding_{i + 1}.0.bias"]))
        hs.append(h)
    g = p["sigma.weight"].expand(xyz.shape[0], -1)  # d sigma / d h8
    g_pe = torch.zeros_like(pe)
    for i in reversed(range(D)):
        g = (g * (hs[i] > 0).to(g.dtype)) @ p[f"xyz_encoding_{i + 1}.0.weight"]
        if i in skips:
            g_pe = g_pe + g[:, : pe.shape[1]]
            g = g[:, pe.shape[1]:]
    g_pe = g_pe + g  # layer 1 input is the PE itself
    gx = g_pe[:, 0:3].clone()
    for k in range(n_freqs):
        f = float(2 ** k)
        s = g_pe[:, 3 + 6 * k: 6 + 6 * k]
        c = g_pe[:, 6 + 6 * k: 9 + 6 * k]
        gx = gx + f * (s * torch.cos(f * xyz) - c * torch.sin(f * xyz))
    return gx


# --------------------------------------------------------------------------------------------
# sampler / compositor
# --------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_importance, det=False, eps=1e-5, u=None, return_inds=False):
    """Inverse-CDF sampling (R/models/rendering.py:7-51). `u` overrides the random draw."""
    n_rays, n_w = weights.shape
    weights = weights + eps
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0, 1, n_importance, device=bins.device).expand(n_rays, n_importance)
        else:
            u = torch.rand(n_rays, n_importance, device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_w)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom[denom < eps] = 1
    samples = bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)
    if return_inds:
        return samples, inds, cdf
    return samples


def composite(results, typ, z_vals, sigmas, rgbs, is_mirrors, normals, pred_normals, *,
              noise_std, white_back, weights_only, mirror_mask=None,
              detach_density_outside_mirror_for_mask_loss=False, detach_density_for_mask_loss=False,
              detach_density_for_normal_loss=False, noise=None):
    """Volume-rendering quadrature of one pass (R/models/rendering.py:175-264)."""
    deltas = z_vals[:, 1:] - z_vals[:, :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    if noise is None:
        noise = torch.randn_like(sigmas)  # drawn even when noise_std == 0 (:189)
    noise = noise * noise_std
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas + noise))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    weights = alphas * torch.cumprod(shifted[:, :-1], -1)
    opacity = weights.sum(1)
    results[f"weights_{typ}"] = weights
    results[f"opacity_{typ}"] = opacity
    results[f"z_vals_{typ}"] = z_vals
    if weights_only:
        return
    rgb_map = (weights.unsqueeze(-1) * rgbs).sum(1)
    depth_map = (weights * z_vals).sum(1)
    if white_back:
        rgb_map += 1 - opacity.unsqueeze(1)
    results[f"rgb_{typ}"] = rgb_map
    results[f"depth_{typ}"] = depth_map
    if is_mirrors is not None:
        if detach_density_for_mask_loss:
            mm = (weights.detach() * is_mirrors).sum(1)
        elif (detach_density_outside_mirror_for_mask_loss and mirror_mask is not None
              and not bool((mirror_mask < 0).any())):
            keep = mirror_mask.clone().bool()
            w2 = weights.clone()
            w2[~keep] = w2[~keep].detach()
            mm = (w2 * is_mirrors).sum(1)
        else:
            mm = (weights * is_mirrors).sum(1)
        results[f"mirror_mask_{typ}"] = mm
    wn = weights.detach() if detach_density_for_normal_loss else weights
    if normals is not None:
        results[f"normal_{typ}"] = normals
        results[f"surface_normal_grad_{typ}"] = (normals * wn.unsqueeze(-1)).sum(1)
    if pred_normals is not None:
        results[f"pred_normal_{typ}"] = pred_normals
        results[f"surface_normal_{typ}"] = (pred_normals * wn.unsqueeze(-1)).sum(1)
    if normals is not None and pred_normals is not None:
        dif = torch.sum((normals - pred_normals) ** 2, dim=-1)
        results[f"normal_dif_{typ}"] = (wn * dif).sum(1)


def coarse_z_vals(rays, n_samples, use_disp=False, perturb=0.0, perturb_u=None):
    """Stratified coarse depths (R/models/rendering.py:271-300)."""
    n_rays = rays.shape[0]
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = torch.linspace(0, 1, n_samples, device=rays.device)
    if not use_disp:
        z = near * (1 - t) + far * t
    else:
        z = 1 / (1 / near * (1 - t) + 1 / far * t)
    z = z.expand(n_rays, n_samples)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        if perturb_u is None:
            perturb_u = torch.rand_like(z)
        z = lower + (upper - lower) * (perturb * perturb_u)
    return z


def render_rays(params, rays, N_samples=64, use_disp=False, perturb=0, noise_std=1, N_importance=0,
                chunk=1024 * 32, white_back=False, test_time=False, *, n_freqs_xyz=10, n_freqs_dir=4,
                rng=None, **kw):
    """One render level (R/models/rendering.py:54-369).

    params: {"coarse": {key: tensor}, "fine": {...}?}.   rng: optional dict with explicit draws
    {"perturb_u","noise_coarse","u_pdf","noise_fine"} (+ the test hook "z_fine", see fine_z); otherwise torch's global RNG is used in the
    reference's call order so that a shared torch.manual_seed gives identical draws.
    """
    rng = rng or {}
    compute_normal = kw.get("compute_normal", True)
    mirror_mask = kw.get("mirror_mask", None)
    flags = dict(
        detach_density_outside_mirror_for_mask_loss=kw.get("detach_density_outside_mirror_for_mask_loss", False),
        detach_density_for_mask_loss=kw.get("detach_density_for_mask_loss", False),
        detach_density_for_normal_loss=kw.get("detach_density_for_normal_loss", False))
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    dir_emb = embed(kw.get("view_dir", rays_d), n_freqs_dir)
    in_dir = dir_emb.shape[1]
    o3, d3 = rays_o.unsqueeze(1), rays_d.unsqueeze(1)
    has_fine = "fine" in params

    def run_pass(results, p, typ, xyz, z_vals, noise):
        S = xyz.shape[1]
        flat = xyz.reshape(-1, 3)
        mm_flat = None if mirror_mask is None else mirror_mask.unsqueeze(-1).repeat(1, S).view(-1)
        dir_flat = dir_emb.unsqueeze(1).expand(n_rays, S, in_dir).reshape(-1, in_dir)
        sig_only = typ == "coarse" and test_time and has_fine
        acc = {k: [] for k in ("sigma", "rgb", "normal", "pred_normal", "is_mirror")}
        for i in range(0, flat.shape[0], chunk):
            pts = flat[i:i + chunk]
            mmc = None if mm_flat is None else mm_flat[i:i + chunk]
            xin = pts if sig_only else torch.cat([pts, dir_flat[i:i + chunk]], 1)
            if "encoder.params" in p:  # nerf_tcnn model family (oracle/hashgrid_oracle.py; identity embeddings)
                from . import hashgrid_oracle as HG
                o = HG.field_forward(p, xin, bound=kw.get("bound", 1.0), sigma_only=sig_only,
                                     compute_normal=compute_normal and not sig_only, mirror_mask=mmc, **flags)
            else:
                o = field_forward(p, xin, n_freqs_xyz=n_freqs_xyz, in_dir=in_dir, compute_normal=compute_normal,
                                  sigma_only=sig_only, mirror_mask=mmc, **flags)
            for k in acc:
                if k in o:
                    acc[k].append(o[k])
        cat = {k: (torch.cat(v, 0) if v else None) for k, v in acc.items()}
        composite(results, typ, z_vals, cat["sigma"].view(n_rays, S),
                  None if cat["rgb"] is None else cat["rgb"].view(n_rays, S, 3),
                  None if cat["is_mirror"] is None else cat["is_mirror"].view(n_rays, S),
                  None if cat["normal"] is None else cat["normal"].view(n_rays, S, 3),
                  None if cat["pred_normal"] is None else cat["pred_normal"].view(n_rays, S, 3),
                  noise_std=noise_std, white_back=white_back, weights_only=sig_only,
                  mirror_mask=mirror_mask, noise=noise, **flags)

    z = coarse_z_vals(rays, N_samples, use_disp, perturb, rng.get("perturb_u"))
    results = {}
    run_pass(results, params["coarse"], "coarse", o3 + d3 * z.unsqueeze(-1), z, rng.get("noise_coarse"))

    def fine_z(z_c):
        if rng.get("z_fine") is not None:
            # test hook: the merged, sorted fine depths of the implementation under test.  The reference detaches sample_pdf's
            # output (R/models/rendering.py:346-349), so no gradient depends on how these depths were obtained; gradient tests of
            # encodings whose Jacobian jumps at cell faces (hash grid) inject them so that both sides evaluate identical positions.
            return rng["z_fine"].to(z_c)
        mid = 0.5 * (z_c[:, :-1] + z_c[:, 1:])
        z_new = sample_pdf(mid, results["weights_coarse"][:, 1:-1].detach(), N_importance,
                           det=(perturb == 0), u=rng.get("u_pdf"))
        return torch.sort(torch.cat([z_c, z_new], -1), -1)[0]

    if N_importance > 0:
        if kw.get("only_one_field", False):
            if kw.get("current_epoch", 0) > kw.get("only_one_field_fine_epoch", 2):
                z = fine_z(z)
                run_pass(results, params["coarse"], "coarse", o3 + d3 * z.unsqueeze(-1), z, rng.get("noise_fine"))
        else:
            z = fine_z(z)
            run_pass(results, params["fine"], "fine", o3 + d3 * z.unsqueeze(-1), z, rng.get("noise_fine"))
    for typ in ("coarse", "fine"):
        if f"depth_{typ}" in results:
            results[f"x_surface_{typ}"] = rays_o + rays_d * results[f"depth_{typ}"].unsqueeze(-1)
    return results


# --------------------------------------------------------------------------------------------
# Whitted recursion (callers of render_rays): R/train.py:129-348, R/eval.py:132-725
# --------------------------------------------------------------------------------------------
def reflect_rays(rays, x_surface, normal, near=0.1):
    """Secondary ray [x_surface, 2(n.w)n - w, 0.1, far] with w = normalize(-d) (R/train.py:219-243)."""
    n = l2_normalize(normal)
    w = l2_normalize(-rays[:, 3:6])
    cos = torch.sum(w * n, dim=-1)
    r = 2 * cos.unsqueeze(-1).repeat(1, 3) * n - w
    return torch.cat([x_surface, r, torch.ones_like(rays[:, 7:8]) * near, rays[:, 7:8]], -1), r


def _threshold_(mask):
    """In-place hard clip of a rendered mirror mask (R/eval.py:305-306, R/train.py:165-166): exactly 0.5 stays."""
    mask[mask > 0.5] = 1
    mask[mask < 0.5] = 0
    return mask


def trace_eval(render_fn, rays, max_recursive_level, level=0, typ="fine", normal_noises=None, trace_ray_times=0,
               noise_fn=None):
    """Eval-semantics recursion, R/eval.py:132-160 (level call), :295-320 (mask + trace condition), :336-360 (normal),
    :506-548 (jitter, reflect, secondary rays, compaction), :609-674 (recursive call + roughness cone), :676-723 (blend).
    Level 0 re-traces ALL rays of a batch that contains a mirror pixel, deeper levels only the mirror rays.

    Roughness cone (--app_control_mirror_roughness): `noise_fn(n) -> (n,3)` returns the ALREADY SCALED normal noise
    (`randn_like(normal) * normal_noise_std`) and is called in the reference's order: once before the first reflection of
    a level (:506-511), then once per extra reflection (:627-631), each draw followed by that reflection's whole sub-tree.
    `normal_noises` (legacy) = a list of tensors consumed in that same call order.  As written the reference adds the
    (N_mirror,3) colour of an extra reflection to the (N_rays,3) colour of the first one at level 0 (:655-666) and so only
    runs when every ray of the batch is a mirror ray; for a partial mask this restatement adds at the mirror rows (the
    evident intent) -- the pinned fixtures use all-mirror batches at level 0, where both coincide."""
    if noise_fn is None and normal_noises is not None:
        it = iter(normal_noises)
        noise_fn = lambda n: next(it)
    res = render_fn(rays)
    res[f"rgb_{typ}_reflect"] = torch.zeros_like(res[f"rgb_{typ}"])
    res[f"depth_{typ}_reflect"] = torch.zeros_like(res[f"depth_{typ}"])
    if f"mirror_mask_{typ}" not in res:
        return res
    mb = _threshold_(res[f"mirror_mask_{typ}"]).bool()
    if bool(mb.any()) and level < max_recursive_level:
        n0 = res[f"surface_normal_{typ}"] if f"surface_normal_{typ}" in res else res[f"surface_normal_grad_{typ}"]
        only_mirror = not (level < 1)
        kw = dict(typ=typ, trace_ray_times=trace_ray_times, noise_fn=noise_fn)
        normal = n0 + noise_fn(n0.shape[0]) if noise_fn is not None else n0
        sec, r = reflect_rays(rays, res[f"x_surface_{typ}"], normal)
        res["reflect_direction"] = r
        if only_mirror:
            sec = sec[mb]
        if sec.shape[0] > 0:
            sub = trace_eval(render_fn, sec, max_recursive_level, level + 1, **kw)
            child = sub[f"rgb_{typ}"]
            if noise_fn is not None:
                for _ in range(trace_ray_times):
                    sec_t, _ = reflect_rays(rays, res[f"x_surface_{typ}"], n0 + noise_fn(n0.shape[0]))
                    sub_t = trace_eval(render_fn, sec_t[mb], max_recursive_level, level + 1, **kw)
                    if only_mirror or sub_t[f"rgb_{typ}"].shape == child.shape:
                        child = child + sub_t[f"rgb_{typ}"]
                    else:
                        child = child.clone()
                        child[mb] = child[mb] + sub_t[f"rgb_{typ}"]
                if only_mirror or bool(mb.all()):
                    child = child / (trace_ray_times + 1)
                else:
                    child = child.clone()
                    child[mb] = child[mb] / (trace_ray_times + 1)
            base = res[f"rgb_{typ}"]
            res[f"rgb_{typ}_direct"] = base
            if only_mirror:
                refl = base.clone()
                refl[mb] = child
                res[f"rgb_{typ}_reflect"][mb] = child
                res[f"depth_{typ}_reflect"][mb] = sub[f"depth_{typ}"]
            else:
                refl = child
                res[f"rgb_{typ}_reflect"] = child
                res[f"depth_{typ}_reflect"] = sub[f"depth_{typ}"]
            m3 = mb.float().unsqueeze(-1).repeat(1, 3)
            res[f"rgb_{typ}"] = m3 * refl + (1 - m3) * base
    return res


def trace_train(render_fn, rays, gt_mirror_mask, max_recursive_level, *, level=0, mask_prev=None, select_type="fine",
                trace_secondary_rays=True, train_geometry_stage=False, only_trace_rays_in_mirrors=True, for_vis=False,
                detach_normal_in_reflection=False, detach_ref_color=False, is_eval=False):
    """Train-semantics recursion, R/train.py:129-348 (`NeRFSystem.render_rays_chunk_recursively`).

    `render_fn(rays) -> dict` is the level call (:132-145: render_rays with the run's hparams and **extra_chunk).
    `gt_mirror_mask` (N,) is extra_chunk["mirror_mask"]: used at level 0 unless it has a negative ("no GT") entry (:155-166);
    it keeps the PARENT length at deeper levels, exactly as the reference passes **extra_chunk down (:254-259).
    `detach_ref_color` = hparams.detach_ref_color_for_blend and current_epoch >= train_geometry_stage_end_epoch + 1 (:276-281)."""
    res = render_fn(rays)
    if mask_prev is None:
        mask_prev = torch.ones(rays.shape[0]).bool()                                   # train.py:116-118
    mirror_mask = gt_mirror_mask.clone()                                               # :155
    if bool((mirror_mask < 0).any()) or level > 0:                                     # :157-166
        if "mirror_mask_fine" in res:
            mirror_mask = res["mirror_mask_fine"].detach()
        elif "mirror_mask_coarse" in res:
            mirror_mask = res["mirror_mask_coarse"].detach()
        else:
            mirror_mask = torch.zeros(rays.shape[0])
        _threshold_(mirror_mask)
    if (not only_trace_rays_in_mirrors) and level > 0:                                 # :167-168
        mirror_mask = mirror_mask * mask_prev.detach()
    mb = mirror_mask.bool()
    trace = trace_secondary_rays and (not train_geometry_stage) and (bool(mb.any()) or for_vis)   # :172-176
    if level >= max_recursive_level:
        trace = False
    t = select_type
    if trace:
        if f"pred_normal_{t}" in res:                                                  # :194-214
            normal = res[f"surface_normal_{t}"] if f"surface_normal_{t}" in res else \
                (res[f"pred_normal_{t}"] * res[f"weights_{t}"].unsqueeze(-1)).sum(1)
        else:
            normal = res[f"surface_normal_grad_{t}"] if f"surface_normal_grad_{t}" in res else \
                (res[f"normal_{t}"] * res[f"weights_{t}"].unsqueeze(-1)).sum(1)
        sec, r = reflect_rays(rays, res[f"x_surface_{t}"], normal.detach() if detach_normal_in_reflection else normal)
        sec_o = res[f"x_surface_{t}"]
        if only_trace_rays_in_mirrors:
            sec = sec[mb]                                                              # :248-252
        if sec.shape[0] > 0:
            sub = trace_train(render_fn, sec, gt_mirror_mask, max_recursive_level, level=level + 1, mask_prev=mirror_mask,
                              select_type=t, trace_secondary_rays=trace_secondary_rays,
                              train_geometry_stage=train_geometry_stage, only_trace_rays_in_mirrors=only_trace_rays_in_mirrors,
                              for_vis=for_vis, detach_normal_in_reflection=detach_normal_in_reflection,
                              detach_ref_color=detach_ref_color, is_eval=is_eval)
            for typ in ("coarse", "fine"):                                             # :263-311
                if f"rgb_{typ}" in res and f"rgb_{typ}" in sub:
                    res[f"rgb_{typ}_direct"] = res[f"rgb_{typ}"]
                    base = res[f"rgb_{typ}"]
                    if only_trace_rays_in_mirrors:
                        refl = base.clone().detach()
                        refl[mb] = sub[f"rgb_{typ}"]
                    else:
                        refl = sub[f"rgb_{typ}"]
                    if detach_ref_color:
                        refl = refl.detach()
                    m3 = mirror_mask.float().unsqueeze(-1).repeat(1, 3)
                    res[f"rgb_{typ}"] = m3 * refl + (1 - m3) * base
                    if is_eval:
                        if only_trace_rays_in_mirrors:
                            res[f"rgb_{typ}_reflect"] = torch.zeros_like(res[f"rgb_{typ}"])
                            res[f"rgb_{typ}_reflect"][mb] = sub[f"rgb_{typ}"]
                        else:
                            res[f"rgb_{typ}_reflect"] = sub[f"rgb_{typ}"]
            if is_eval:                                                                # :312-326
                if only_trace_rays_in_mirrors:
                    res[f"depth_{t}_reflect"] = torch.zeros_like(res[f"depth_{t}"])
                    res[f"depth_{t}_reflect"][mb] = sub[f"depth_{t}"]
                else:
                    res[f"depth_{t}_reflect"] = sub[f"depth_{t}"]
                res["secondary_rays_o"] = sec_o
                res["reflect_direction"] = r
    elif is_eval:                                                                      # :327-346
        for typ in ("coarse", "fine"):
            if f"rgb_{typ}" in res:
                res[f"rgb_{typ}_reflect"] = torch.zeros_like(res[f"rgb_{typ}"])
                res[f"rgb_{typ}_direct"] = torch.zeros_like(res[f"rgb_{typ}"])
        res[f"depth_{t}_reflect"] = torch.zeros_like(res[f"depth_{t}"])
        res["secondary_rays_o"] = torch.zeros_like(res[f"rgb_{t}"])
        res["reflect_direction"] = torch.zeros_like(res[f"rgb_{t}"])
    return res
