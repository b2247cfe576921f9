"""Drop-in for R/models/rendering.py: ``render_rays(models, embeddings, rays, N_samples, use_disp, perturb,
noise_std, N_importance, chunk, white_back, test_time=False, **kwargs) -> dict`` with the reference's
positional signature (R/models/rendering.py:54-67) and output-key contract, executed by the
sm_100a kernels of libmnrf.so.  ``sample_pdf`` (rendering.py:7-51) is exposed the same way.

Differences from the reference that a caller can observe:
  * inputs must be CUDA float32 tensors -- there is no CPU path (RuntimeError otherwise);
  * ``chunk`` does not change results (the reference's don't depend on it either); it is ignored, the ray batch
    is processed in sub-batches sized to bound scratch memory;
  * with autograd enabled and parameters that require grad, the level runs through the training path
    (``autograd.py`` / ``csrc/train.cu``: fp32 layer-by-layer kernels with a hand-written backward); gradients reach the
    parameters of both models and, when the rays require grad, the rays' origins and directions;
  * a batch of exactly one ray works (the reference's ``.squeeze()`` at rendering.py:365 breaks it);
  * ``view_dir`` (rendering.py:276; never passed by any reference caller) is honoured by the inference path and refused
    under autograd.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from .mirror_nerf import _ptr, _stream_ptr, packed_field

__all__ = ["render_rays", "sample_pdf"]

# rays per internal launch group: bounds scratch to ~ MAX_RAYS * (S+Ni) * 32 B
MAX_RAYS_PER_LAUNCH = int(os.environ.get("MNRF_MAX_RAYS_PER_LAUNCH", 1 << 17))
DEFAULT_IMPL = os.environ.get("MNRF_FIELD_IMPL", "tc3")  # tc3 | tc1 | fp32

_tables = {}


def _linspace(n, device):
    """torch.linspace(0,1,n) -- taken from torch so the values are bit-identical to the reference's
     and cached per device."""
    key = (n, str(device))
    t = _tables.get(key)
    if t is None:
        t = torch.linspace(0, 1, n, device=device, dtype=torch.float32)
        _tables[key] = t
    return t


def _check_rays(rays):
    if not isinstance(rays, torch.Tensor) or rays.dim() != 2 or rays.shape[1] != 8:
        raise RuntimeError(f"rays must be a (N,8) tensor [o,d,near,far], got {getattr(rays, 'shape', None)}")
    if not rays.is_cuda:
        raise RuntimeError("rays must be a CUDA tensor: the B200 renderer has no CPU fallback")
    if rays.dtype != torch.float32:
        raise RuntimeError(f"rays must be float32, got {rays.dtype}")
    return rays.detach().contiguous()


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, return_inds=False):
    """Inverse-CDF sampling (R/models/rendering.py:7-51) on CUDA.  bins (N, n_w+1), weights (N, n_w).
    Returns samples (N, N_importance) [, inds int64, cdf]."""
    if eps != 1e-5:
        raise NotImplementedError("sample_pdf: eps is fixed to the reference's 1e-5")
    if not (bins.is_cuda and weights.is_cuda):
        raise RuntimeError("sample_pdf: CUDA tensors required (no CPU path)")
    lib = _lib.load()
    n, nw = weights.shape
    dev = weights.device
    bins = bins.detach().contiguous().float()
    weights = weights.detach().contiguous().float()
    if bins.shape != (n, nw + 1):
        raise RuntimeError(f"sample_pdf: bins must be (N, n_w+1), got {tuple(bins.shape)} for weights {tuple(weights.shape)}")
    u = _linspace(N_importance, dev) if det else torch.rand(n, N_importance, device=dev)
    samples = torch.empty(n, N_importance, device=dev, dtype=torch.float32)
    inds = torch.empty(n, N_importance, device=dev, dtype=torch.int64) if return_inds else None
    cdf = torch.empty(n, nw + 1, device=dev, dtype=torch.float32) if return_inds else None
    with torch.cuda.device(dev):
        _lib.check(lib.mnrf_sample_pdf_bins(_ptr(bins), _ptr(weights), n, nw, N_importance, _ptr(u),
                                            0 if det else N_importance, _ptr(samples), _ptr(inds), _ptr(cdf),
                                            _stream_ptr()), "mnrf_sample_pdf_bins")
    if return_inds:
        return samples, inds, cdf
    return samples


def _render_level_train(lib, models, rays, Sc, Ni, second_pass, rerun, sig_only, use_disp, perturb, noise_std, white_back,
                        compute_normal, kwargs, perturb_u, noise_c, u_pdf, noise_f):
    """render_rays with gradients (rendering.py:268-369 under autograd): per group of rays, coarse depths -> differentiable
    coarse pass -> inverse-CDF resampling on the DETACHED coarse weights (rendering.py:335,353) -> differentiable second
    pass.  Each pass is one autograd node (autograd._PassFn)."""
    from . import autograd as AG
    dev = rays.device
    n = rays.shape[0]
    Sf = Sc + Ni
    detach_mask = bool(kwargs.get("detach_density_for_mask_loss", False))
    detach_normal = bool(kwargs.get("detach_density_for_normal_loss", False))
    ray_detach_all = None
    mm = kwargs.get("mirror_mask")
    if kwargs.get("detach_density_outside_mirror_for_mask_loss", False) and mm is not None and not detach_mask:
        mm = mm.to(dev).reshape(-1)
        if not bool((mm < 0).any()):  # rendering.py:227-231 (host sync, as in the reference)
            if mm.shape[0] != n:
                raise RuntimeError(f"mirror_mask has {mm.shape[0]} entries for {n} rays")
            ray_detach_all = (~mm.bool()).float().contiguous()
    z_steps = _linspace(Sc, dev)
    u_det = _linspace(Ni, dev) if Ni > 0 else None
    second_module = models["coarse"] if rerun else models.get("fine")
    parts_c, parts_f = [], []
    cut = lambda t, lo, hi: None if t is None else t[lo:hi].contiguous()
    for lo in range(0, max(n, 1), AG.MAX_TRAIN_RAYS):
        hi = min(n, lo + AG.MAX_TRAIN_RAYS)
        m = hi - lo
        r = rays[lo:hi].contiguous()
        rd = cut(ray_detach_all, lo, hi)
        z_c = torch.empty(m, Sc, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.mnrf_coarse_z(_ptr(r), m, _ptr(z_steps), Sc, int(bool(use_disp)), float(perturb),
                                         _ptr(cut(perturb_u, lo, hi)), _ptr(z_c), _stream_ptr()), "mnrf_coarse_z")
        common = dict(compute_normal=compute_normal, white_back=bool(white_back), noise_std=float(noise_std),
                      detach_mask=detach_mask, detach_normal=detach_normal)
        oc = AG.run_pass(models["coarse"], r, z_c, cut(noise_c, lo, hi), rd, **common)
        oc["z_vals"] = z_c
        if sig_only:  # rendering.py:208-209: the coarse pass only publishes weights / opacity / z_vals at test time
            oc = {k: oc[k] for k in ("weights", "opacity", "z_vals")}
        parts_c.append(oc)
        if second_pass:
            z_f = torch.empty(m, Sf, device=dev, dtype=torch.float32)
            u = cut(u_pdf, lo, hi) if u_pdf is not None else u_det
            with torch.cuda.device(dev):
                _lib.check(lib.mnrf_sample_pdf(_ptr(z_c), _ptr(oc["weights"].detach()), m, Sc, Ni, _ptr(u),
                                               Ni if u_pdf is not None else 0, _ptr(z_f), None, None, None,
                                               _stream_ptr()), "mnrf_sample_pdf")
            of = AG.run_pass(second_module, r, z_f, cut(noise_f, lo, hi), rd, **common)
            of["z_vals"] = z_f
            parts_f.append(of)

    def merge(parts):
        return {k: (parts[0][k] if len(parts) == 1 else torch.cat([p[k] for p in parts], 0)) for k in parts[0]}

    order = ("weights", "opacity", "z_vals", "rgb", "depth", "mirror_mask", "normal", "surface_normal_grad",
             "pred_normal", "surface_normal", "normal_dif")
    res = {}
    tc = merge(parts_c)
    for k in order:
        if k in tc:
            res[f"{k}_coarse"] = tc[k]
    if second_pass:
        tf = merge(parts_f)
        typ = "coarse" if rerun else "fine"
        for k in order:
            if k in tf:
                res[f"{k}_{typ}"] = tf[k]
        if "x_surface" in tc and not rerun:
            res["x_surface_coarse"] = tc["x_surface"]
        res[f"x_surface_{typ}"] = tf["x_surface"]
    elif "x_surface" in tc:
        res["x_surface_coarse"] = tc["x_surface"]
    return res


def render_rays(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=1, N_importance=0,
                chunk=1024 * 32, white_back=False, test_time=False, **kwargs):
    """One render level.  See the module docstring; kwargs consumed: compute_normal (default True, like the
    reference), only_one_field, current_epoch, only_one_field_fine_epoch, field_impl ("tc3" | "tc1" | "fp32"),
    rng (dict of explicit draws: perturb_u, noise_coarse, u_pdf, noise_fine -- for tests).  mirror_mask and the three
    detach_* flags only shape gradients: they are honoured by the training path and have no effect on forward values."""
    lib = _lib.load()
    rays_in = rays
    rays = _check_rays(rays)
    dev = rays.device
    n = rays.shape[0]
    if "coarse" not in models:
        raise KeyError("models must contain 'coarse'")
    from .mirror_nerf_tcnn import is_hash_field, packed_hash_field
    hashed = is_hash_field(models["coarse"])  # nerf_tcnn model family (BASELINE config 3): identity embeddings, [xyz | d] input
    for name, want in (("xyz", 0 if hashed else 10), ("dir", 0 if hashed else 4)):
        emb = embeddings.get(name) if isinstance(embeddings, dict) else None
        nf = getattr(emb, "N_freqs", want)
        if nf != want:
            raise NotImplementedError(f"embedding_{name}.N_freqs={nf}: the {'hash-grid' if hashed else 'MLP'} field takes "
                                      f"N_freqs={want} (R/train.py:45-47,69-70)")
    pack = packed_hash_field if hashed else packed_field
    view_dir = kwargs.get("view_dir")
    dir_source = None
    if view_dir is not None:  # rendering.py:276: embedding_dir(kwargs.get("view_dir", rays_d)); the ray geometry keeps rays_d
        if hashed:
            raise NotImplementedError("view_dir with the hash-grid field is not built")
        view_dir = view_dir.detach().to(device=dev, dtype=torch.float32)
        if tuple(view_dir.shape) != (n, 3):
            raise RuntimeError(f"view_dir must be (N,3), got {tuple(view_dir.shape)}")
        dir_source = rays.clone()
        dir_source[:, 3:6] = view_dir
    if "coarse" not in models:
        raise KeyError("models must contain 'coarse'")
    compute_normal = bool(kwargs.get("compute_normal", True))
    only_one_field = bool(kwargs.get("only_one_field", False))
    rerun = (N_importance > 0 and only_one_field
             and kwargs.get("current_epoch", 0) > kwargs.get("only_one_field_fine_epoch", 2))
    has_fine_model = "fine" in models
    second_pass = N_importance > 0 and (rerun or (has_fine_model and not only_one_field))
    if N_importance > 0 and not only_one_field and not has_fine_model:
        raise KeyError("N_importance > 0 needs models['fine'] (or only_one_field=True)")
    params = list(models["coarse"].parameters()) if hasattr(models["coarse"], "parameters") else []
    if has_fine_model and hasattr(models["fine"], "parameters"):
        params += list(models["fine"].parameters())
    needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or rays_in.requires_grad)

    coarse = pack(models["coarse"])
    fine = pack(models["fine"]) if (has_fine_model and not only_one_field) else None
    # NB rendering.py:139 tests '"fine" in models' for the sigma-only shortcut
    sig_only = bool(test_time) and has_fine_model
    fine_for_lib = fine if fine is not None else (coarse if sig_only else None)
    impl = _lib.IMPL_BY_NAME[kwargs.get("field_impl", DEFAULT_IMPL)]
    Sc, Ni = int(N_samples), int(N_importance) if second_pass else 0
    Sf = Sc + Ni
    rng = kwargs.get("rng") or {}

    def draw(name, shape, fn):
        t = rng.get(name)
        if t is None:
            t = fn(shape, device=dev, dtype=torch.float32)
        else:
            t = t.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"rng['{name}'] has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t

    # RNG tensors come from torch, drawn in the reference's call order: rand_like(z_vals) if perturb > 0
    # (rendering.py:298), randn_like(sigmas) in EVERY inference() call even when noise_std == 0 (:189), rand(N, N_importance) in
    # sample_pdf if perturb != 0 (:29-31).  The noise_std == 0 draws are made and discarded so that a seeded run consumes the
    # same generator stream as the reference on the same device (consume_rng=False skips them).
    consume = bool(kwargs.get("consume_rng", True))
    perturb_u = draw("perturb_u", (n, Sc), torch.rand) if perturb > 0 else None
    noise_c = draw("noise_coarse", (n, Sc), torch.randn) if (noise_std != 0 or consume) else None
    u_pdf = draw("u_pdf", (n, Ni), torch.rand) if (second_pass and perturb != 0) else None
    noise_f = draw("noise_fine", (n, Sf), torch.randn) if (second_pass and (noise_std != 0 or consume)) else None
    if noise_std == 0:
        noise_c = noise_f = None

    if needs_grad:
        if dir_source is not None:
            raise NotImplementedError("view_dir is not supported under autograd (no reference caller passes it)")
        return _render_level_train(lib, models, rays_in.contiguous(), Sc, Ni, second_pass, rerun, sig_only, use_disp, perturb, noise_std,
                                   white_back, compute_normal, kwargs, perturb_u, noise_c, u_pdf, noise_f)

    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    res = {}

    def alloc_pass(S, full, field):
        t = {"weights": new(n, S), "opacity": new(n), "z_vals": new(n, S)}
        if full:
            t.update(rgb=new(n, 3), depth=new(n), x_surface=new(n, 3))
            if field.has_mirror:
                t["mirror_mask"] = new(n)
            if compute_normal:
                t["normal"] = new(n, S, 3)
                t["surface_normal_grad"] = new(n, 3)
            if field.has_normal:
                t["pred_normal"] = new(n, S, 3)
                t["surface_normal"] = new(n, 3)
                if compute_normal:
                    t["normal_dif"] = new(n)
        return t

    tc = alloc_pass(Sc, not sig_only, coarse)
    tf = alloc_pass(Sf, True, coarse if rerun else fine) if second_pass else None

    cfg = _lib.LevelCfg(n_samples=Sc, n_importance=Ni, use_disp=int(bool(use_disp)), perturb=float(perturb),
                        noise_std=float(noise_std), white_back=int(bool(white_back)), test_time=int(sig_only),
                        compute_normal=int(compute_normal), rerun_coarse_on_fine=int(rerun), impl=impl,
                        early_termination_eps=0.0, no_fused_composite=int(not kwargs.get("fused_composite", True)),
                        dir_source=None, stats=None)
    z_steps = _linspace(Sc, dev)
    u_det = _linspace(Ni, dev) if Ni > 0 else None

    def comp_struct(t, lo, hi, S):
        def p(key, per_ray):
            x = t.get(key) if t is not None else None
            return None if x is None else C.c_void_p(x.data_ptr() + lo * per_ray * 4)
        return _lib.CompositeOut(weights=p("weights", S), opacity=p("opacity", 1), rgb=p("rgb", 3), depth=p("depth", 1),
                                 mirror_mask=p("mirror_mask", 1), pred_normal=p("pred_normal", 3 * S),
                                 surface_normal=p("surface_normal", 3), surface_normal_grad=p("surface_normal_grad", 3),
                                 normal_dif=p("normal_dif", 1), x_surface=p("x_surface", 3))

    def off(x, lo, per_ray):
        return None if x is None else C.c_void_p(x.data_ptr() + lo * per_ray * 4)

    with torch.cuda.device(dev):
        stream = _stream_ptr()
        step = max(1, min(MAX_RAYS_PER_LAUNCH, ((1 << 31) - 1) // max(Sf, 1) // 2))
        ws = None
        for lo in range(0, n, step):
            m = min(step, n - lo)
            need = int(lib.mnrf_level_workspace_bytes_for(coarse.handle, None if fine_for_lib is None else fine_for_lib.handle,
                                                          m, C.byref(cfg), int(noise_c is not None or noise_f is not None)))
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, device=dev, dtype=torch.uint8)
            out = _lib.LevelOut(
                z_coarse=off(tc["z_vals"], lo, Sc), coarse=comp_struct(tc, lo, lo + m, Sc),
                normal_coarse=off(tc.get("normal"), lo, 3 * Sc),
                z_fine=off(tf["z_vals"], lo, Sf) if tf else None, fine=comp_struct(tf, lo, lo + m, Sf),
                normal_fine=off(tf.get("normal"), lo, 3 * Sf) if tf else None)
            cfg.dir_source = None if dir_source is None else dir_source.data_ptr() + lo * 8 * 4
            rs = _lib.LevelRng(perturb_u=off(perturb_u, lo, Sc), noise_coarse=off(noise_c, lo, Sc),
                               u_pdf=off(u_pdf, lo, Ni), noise_fine=off(noise_f, lo, Sf))
            _lib.check(lib.mnrf_render_level(
                coarse.handle, None if fine_for_lib is None else fine_for_lib.handle, off(rays, lo, 8), m,
                C.byref(cfg), C.byref(rs), _ptr(z_steps), _ptr(u_det), _ptr(ws), need, C.byref(out), stream),
                "mnrf_render_level")

    order = ("weights", "opacity", "z_vals", "rgb", "depth", "mirror_mask", "normal", "surface_normal_grad",
             "pred_normal", "surface_normal", "normal_dif")

    def publish(t, typ):
        for k in order:
            if k in t:
                res[f"{k}_{typ}"] = t[k]

    publish(tc, "coarse")
    if second_pass:
        publish(tf, "coarse" if rerun else "fine")  # only_one_field overwrites the coarse keys (rendering.py:328-348)
    if "x_surface" in tc and not (second_pass and rerun):
        res["x_surface_coarse"] = tc["x_surface"]
    if second_pass:
        res["x_surface_coarse" if rerun else "x_surface_fine"] = tf["x_surface"]
    return res
