"""Whitted-style recursion around render_rays -- SURVEY.md section 8f row 1 -- restating what the reference's callers do
(R/eval.py:132-160,295-320,515-548,676-697; train semantics R/train.py:153-296) with the mask threshold, reflection,
stable compaction and blend running as CUDA kernels of libmnrf.so instead of boolean-mask indexing.

This is an ADDITIVE entry point: ``render_rays`` itself stays single-level and drop-in (the reference puts the bounce
in its callers).
"""
from __future__ import annotations

import torch

from . import _lib
from .mirror_nerf import _ptr, _stream_ptr
from .rendering import render_rays

__all__ = ["render_rays_recursive", "render_rays_recursive_device", "reflect_rays", "compact_rows", "blend_reflection", "axpy_rows"]

RAY_FORWARD_OFFSET = 0.1  # near of a secondary ray (eval.py:529, train.py:232)


def reflect_rays(rays, x_surface, normal, mask):
    """Threshold `mask` IN PLACE (>0.5 -> 1, <0.5 -> 0) and build secondary rays.
    Returns (secondary (n,8), reflect_direction (n,3), any_mirror: 0-d int32 device tensor)."""
    lib = _lib.load()
    n = rays.shape[0]
    dev = rays.device
    sec = torch.empty(n, 8, device=dev, dtype=torch.float32)
    refl = torch.empty(n, 3, device=dev, dtype=torch.float32)
    flag = torch.zeros((), device=dev, dtype=torch.int32)
    for t in (rays, x_surface, normal, mask):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.mnrf_reflect_rays(_ptr(rays), _ptr(x_surface), _ptr(normal), _ptr(mask), n, RAY_FORWARD_OFFSET,
                                         _ptr(sec), _ptr(refl), _ptr(flag), _stream_ptr()), "mnrf_reflect_rays")
    return sec, refl, flag


def compact_rows(rows, mask):
    """rows[mask != 0] in the same (stable) order as boolean indexing.  Returns (compacted, index int32 (n,) with the
    destination row or -1).  Reads the count back to size the result (one host sync, like `.any()` in the reference)."""
    lib = _lib.load()
    n, c = rows.shape
    dev = rows.device
    out = torch.empty(n, c, device=dev, dtype=torch.float32)
    index = torch.empty(n, device=dev, dtype=torch.int32)
    count = torch.zeros((), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(lib.mnrf_compact_rows(_ptr(rows), _ptr(mask), n, c, _ptr(out), _ptr(index), _ptr(count),
                                         _stream_ptr()), "mnrf_compact_rows")
    return out[: int(count.item())], index


def blend_reflection(base_rgb, mask, child_rgb, child_depth, index=None):
    """rgb = m * reflect + (1-m) * base with m = (mask != 0).  Returns (rgb, rgb_reflect, depth_reflect)."""
    lib = _lib.load()
    n = base_rgb.shape[0]
    dev = base_rgb.device
    rgb = torch.empty(n, 3, device=dev, dtype=torch.float32)
    rgb_reflect = torch.empty(n, 3, device=dev, dtype=torch.float32)
    depth_reflect = torch.empty(n, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.mnrf_blend_reflection(_ptr(base_rgb), _ptr(mask), _ptr(child_rgb.contiguous()),
                                             _ptr(child_depth.contiguous()), _ptr(index), n, _ptr(rgb), _ptr(rgb_reflect),
                                             _ptr(depth_reflect), _stream_ptr()), "mnrf_blend_reflection")
    return rgb, rgb_reflect, depth_reflect


def axpy_rows(dense, compact, index, alpha, beta):
    """dense[i] = alpha*dense[i] + beta*compact[index[i]] for rows with index[i] >= 0, in place."""
    lib = _lib.load()
    n, c = dense.shape
    with torch.cuda.device(dense.device):
        _lib.check(lib.mnrf_axpy_rows(_ptr(dense), _ptr(None if compact is None else compact.contiguous()), _ptr(index),
                                      n, c, float(alpha), float(beta), _stream_ptr()), "mnrf_axpy_rows")
    return dense


def render_rays_recursive(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=0,
                          N_importance=0, chunk=1024 * 32, white_back=False, max_recursive_level=1,
                          only_trace_rays_in_mirrors=None, test_time=True, _level=0, render_fn=None,
                          normal_noise_std=0.0, trace_ray_times=0, normal_noises=None, compact_outputs=False, **kwargs):
    """Eval-semantics recursion (R/eval.py:132-725): level 0 re-traces ALL rays of a batch that contains a mirror
    pixel, deeper levels only the mirror rays; `only_trace_rays_in_mirrors=True` compacts at every level (train.py
    semantics).  Returns the level-0 render_rays dict with rgb_{typ} blended plus rgb_{typ}_direct/_reflect,
    depth_{typ}_reflect and reflect_direction.  `render_fn(rays) -> dict` replaces render_rays (used by the tests to
    check the recursion logic on a smooth stand-in field).

    Roughness cone (R/eval.py:506-511,623-674, --app_control_mirror_roughness): with
    `normal_noise_std` > 0 the surface normal is jittered by N(0, std^2) before reflecting, and `trace_ray_times` extra
    jittered reflections of the MIRROR rays are rendered and averaged with the first one (the reference adds an
    (N_mirror,3) tensor to an (N_rays,3) one at level 0 -- a latent shape bug; the evident intent, mirror rays only, is
    what is implemented).  `normal_noises`: optional list of trace_ray_times+1 explicit (n,3) noise tensors (tests)."""
    kwargs.setdefault("compute_normal", False)
    if compact_outputs:
        if render_fn is not None or _level != 0:
            raise ValueError("compact_outputs=True is the device-side recursion of libmnrf.so: it renders the real field")
        return render_rays_recursive_device(models, embeddings, rays, N_samples, use_disp, perturb, noise_std, N_importance,
                                            white_back, max_recursive_level, only_trace_rays_in_mirrors, test_time,
                                            normal_noise_std, trace_ray_times, normal_noises, **kwargs)
    if render_fn is None and torch.is_grad_enabled() and any(
            p.requires_grad for m in models.values() if hasattr(m, "parameters") for p in m.parameters()):
        # the reflect / compact / blend kernels of this driver are not differentiable: refuse instead of silently cutting the graph
        raise NotImplementedError(
            "render_rays_recursive is the inference (eval-semantics) driver; for training call render_rays per level and build "
            "the secondary rays / blend with torch ops as R/train.py:194-296 does (render_rays is differentiable w.r.t. "
            "parameters and rays), or wrap this call in torch.no_grad()")
    if render_fn is not None:
        res = render_fn(rays)
    else:
        res = render_rays(models, embeddings, rays, N_samples, use_disp, perturb, noise_std, N_importance, chunk,
                          white_back, test_time=test_time, **kwargs)
    typ = "fine" if (N_importance > 0 and not kwargs.get("only_one_field", False)) else "coarse"
    if f"mirror_mask_{typ}" not in res:
        return res
    rays = rays.detach().contiguous()
    normal = res.get(f"surface_normal_{typ}", res.get(f"surface_normal_grad_{typ}"))
    if normal is None:
        raise RuntimeError("render_rays_recursive: the field has no normal_net, so the reflection needs the analytic normal: "
                           "pass compute_normal=True (R/eval.py:146-148,351-360)")
    mask = res[f"mirror_mask_{typ}"]
    rough = normal_noise_std > 0

    def jitter(t):
        if not rough:
            return normal
        nz = normal_noises[t].to(normal) if normal_noises is not None else torch.randn_like(normal) * normal_noise_std
        return (normal + nz).contiguous()

    sec, refl, flag = reflect_rays(rays, res[f"x_surface_{typ}"], jitter(0), mask)
    base = res[f"rgb_{typ}"]
    res[f"rgb_{typ}_reflect"] = torch.zeros_like(base)
    res[f"depth_{typ}_reflect"] = torch.zeros_like(res[f"depth_{typ}"])
    if _level >= max_recursive_level or int(flag.item()) == 0:  # the reference's mirror_mask.any() host sync
        return res
    res["reflect_direction"] = refl
    only_mirror = (_level >= 1) if only_trace_rays_in_mirrors is None else bool(only_trace_rays_in_mirrors)
    index = None
    if only_mirror:
        sec, index = compact_rows(sec, mask)
    if sec.shape[0] == 0:
        return res
    sub = render_rays_recursive(models, embeddings, sec, N_samples, use_disp, perturb, noise_std, N_importance, chunk,
                                white_back, max_recursive_level, only_trace_rays_in_mirrors, test_time,
                                _level=_level + 1, render_fn=render_fn, normal_noise_std=normal_noise_std,
                                trace_ray_times=trace_ray_times, **kwargs)
    if rough and trace_ray_times > 0:
        # extra jittered reflections of the mirror rays only, averaged into the first one
        acc = sub[f"rgb_{typ}"].clone()
        for t in range(1, trace_ray_times + 1):
            sec_t, _, _ = reflect_rays(rays, res[f"x_surface_{typ}"], jitter(t), mask)
            sec_m, index_m = compact_rows(sec_t, mask)
            if sec_m.shape[0] == 0:
                break
            sub_t = render_rays_recursive(models, embeddings, sec_m, N_samples, use_disp, perturb, noise_std,
                                          N_importance, chunk, white_back, max_recursive_level,
                                          only_trace_rays_in_mirrors, test_time, _level=_level + 1, render_fn=render_fn,
                                          normal_noise_std=normal_noise_std, trace_ray_times=trace_ray_times, **kwargs)
            if only_mirror:   # acc is compacted too, same order
                axpy_rows(acc, sub_t[f"rgb_{typ}"], None, 1.0, 1.0)
            else:             # acc is dense: add at the mirror rows
                axpy_rows(acc, sub_t[f"rgb_{typ}"], index_m, 1.0, 1.0)
        if only_mirror:
            axpy_rows(acc, None, None, 1.0 / (trace_ray_times + 1), 0.0)
        else:
            _, index_m = compact_rows(sec, mask)
            axpy_rows(acc, None, index_m, 1.0 / (trace_ray_times + 1), 0.0)
        sub = dict(sub)
        sub[f"rgb_{typ}"] = acc
    rgb, rgb_reflect, depth_reflect = blend_reflection(base, mask, sub[f"rgb_{typ}"], sub[f"depth_{typ}"], index)
    res[f"rgb_{typ}_direct"] = base
    res[f"rgb_{typ}"] = rgb
    res[f"rgb_{typ}_reflect"] = rgb_reflect
    res[f"depth_{typ}_reflect"] = depth_reflect
    return res


def render_rays_recursive_device(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=0, N_importance=0,
                                 white_back=False, max_recursive_level=1, only_trace_rays_in_mirrors=None, test_time=True,
                                 normal_noise_std=0.0, trace_ray_times=0, normal_noises=None, noise_seed=0,
                                 workspace_budget_bytes=None, with_level_rays=False, early_termination_eps=1e-5,
                                 with_stats=False, **kwargs):
    """The whole eval-semantics recursion (R/eval.py::batched_inference :114-740, incl. --app_control_mirror_roughness) as ONE
    call of ``mnrf_render_recursive``: no host synchronisation between level 0 and the blend, mirror rays counted / compacted /
    re-enqueued on the device, the T+1 jittered reflections of a level rendered as one child batch.

    Returns the compact per-ray dict of the select type ``t`` (fine, or coarse without a second pass):
    ``rgb_t`` (blended), ``rgb_t_direct``, ``rgb_t_reflect``, ``depth_t``, ``depth_t_reflect``, ``opacity_t``, ``mirror_mask_t``
    (hard-clipped, as eval.py:305-306 leaves it), ``surface_normal_t`` / ``surface_normal_grad_t``, ``x_surface_t``,
    ``reflect_direction`` -- the per-sample tensors (weights, z_vals, pred_normal) are not materialised on this path.
    ``normal_noises``: optional list of trace_ray_times+1 (n,3) standard-normal*std tensors for the LEVEL-0 reflections.
    ``early_termination_eps``: the fused fine pass stops a ray once its transmittance is below this (every skipped sample has
    weight < eps, so |d rgb| < eps; 0 = composite all samples as the reference does).  ``with_stats``: adds ``fused_stats`` =
    int64 [tiles executed, 32-sample chunks skipped] of all fine passes of the call."""
    import ctypes as C

    from .rendering import DEFAULT_IMPL, _check_rays, _linspace
    from .mirror_nerf import packed_field
    from .mirror_nerf_tcnn import is_hash_field, packed_hash_field
    if perturb != 0 or noise_std != 0:
        raise NotImplementedError("render_rays_recursive_device renders with eval semantics (perturb = noise_std = 0, eval.py:141-142)")
    if only_trace_rays_in_mirrors is False:
        raise NotImplementedError("only_trace_rays_in_mirrors=False at every level (train.py's mask *= mask_prev) is not built; "
                                  "None = eval.py semantics, True = compact at every level")
    lib = _lib.load()
    rays = _check_rays(rays)
    dev = rays.device
    n = rays.shape[0]
    hashed = is_hash_field(models["coarse"])
    pack = packed_hash_field if hashed else packed_field
    only_one_field = bool(kwargs.get("only_one_field", False))
    has_fine = "fine" in models
    rerun = (N_importance > 0 and only_one_field
             and kwargs.get("current_epoch", 0) > kwargs.get("only_one_field_fine_epoch", 2))
    second = N_importance > 0 and (rerun or (has_fine and not only_one_field))
    coarse = pack(models["coarse"])
    fine = pack(models["fine"]) if (has_fine and not only_one_field) else None
    sig_only = bool(test_time) and has_fine
    fine_for_lib = fine if fine is not None else (coarse if sig_only else None)
    last = (coarse if rerun else fine) if second else coarse
    typ = "coarse" if (rerun or not second) else "fine"
    Sc, Ni = int(N_samples), int(N_importance) if second else 0
    impl = _lib.IMPL_BY_NAME[kwargs.get("field_impl", DEFAULT_IMPL)]
    compute_normal = bool(kwargs.get("compute_normal", False)) and not last.has_normal
    lc = _lib.LevelCfg(n_samples=Sc, n_importance=Ni, use_disp=int(bool(use_disp)), perturb=0.0, noise_std=0.0,
                       white_back=int(bool(white_back)), test_time=int(sig_only), compute_normal=int(compute_normal),
                       rerun_coarse_on_fine=int(rerun), impl=impl, early_termination_eps=float(early_termination_eps or 0.0),
                       no_fused_composite=int(not kwargs.get("fused_composite", True)), dir_source=None, stats=None)
    stats = torch.zeros(2, device=dev, dtype=torch.int64) if with_stats else None
    if stats is not None:
        lc.stats = stats.data_ptr()
    T = int(trace_ray_times) if normal_noise_std > 0 else 0
    cfg = _lib.TraceCfg(level=lc, max_recursive_level=int(max_recursive_level),
                        only_trace_rays_in_mirrors=1 if only_trace_rays_in_mirrors else -1, trace_ray_times=T,
                        normal_noise_std=float(normal_noise_std), noise_seed=int(noise_seed))
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    t = dict(rgb=new(n, 3), rgb_direct=new(n, 3), rgb_reflect=new(n, 3), depth=new(n), depth_reflect=new(n), opacity=new(n),
             surface_normal=new(n, 3), x_surface=new(n, 3), reflect_direction=new(n, 3))
    if last.has_mirror:
        t["mirror_mask"] = new(n)
    level_rays = torch.zeros(int(max_recursive_level) + 1, device=dev, dtype=torch.int32) if with_level_rays else None
    noise0 = None
    if normal_noises is not None and T >= 0 and normal_noise_std > 0:
        if len(normal_noises) != T + 1:
            raise ValueError(f"normal_noises must hold trace_ray_times+1 = {T + 1} tensors")
        # the C side multiplies standard-normal draws by the std; explicit tensors are already scaled -> divide it out exactly
        # when std is a power of two, otherwise accept the 1-ulp difference
        noise0 = (torch.stack([z.to(device=dev, dtype=torch.float32) for z in normal_noises], 0) / float(normal_noise_std)).contiguous()
        if tuple(noise0.shape) != (T + 1, n, 3):
            raise ValueError(f"normal_noises tensors must be (n,3); got {tuple(noise0.shape)}")
    with torch.cuda.device(dev):
        budget = int(workspace_budget_bytes) if workspace_budget_bytes else 0
        if budget == 0 and T > 0:
            free, _total = torch.cuda.mem_get_info(dev)
            budget = int(free * 0.6)
        need = int(lib.mnrf_recursive_workspace_bytes(coarse.handle, None if fine_for_lib is None else fine_for_lib.handle, n,
                                                      C.byref(cfg), budget))
        if need < 0:
            _lib.check(1, "mnrf_recursive_workspace_bytes")
        ws = torch.empty(max(need, 256), device=dev, dtype=torch.uint8)
        out = _lib.TraceOut(**{k: _ptr(v) for k, v in t.items()}, level_rays=_ptr(level_rays))
        _lib.check(lib.mnrf_render_recursive(coarse.handle, None if fine_for_lib is None else fine_for_lib.handle, _ptr(rays), n,
                                             C.byref(cfg), _ptr(_linspace(Sc, dev)), _ptr(_linspace(Ni, dev)) if Ni > 0 else None,
                                             _ptr(noise0), _ptr(ws), need, C.byref(out), _stream_ptr()), "mnrf_render_recursive")
    res = {f"rgb_{typ}": t["rgb"], f"rgb_{typ}_direct": t["rgb_direct"], f"rgb_{typ}_reflect": t["rgb_reflect"],
           f"depth_{typ}": t["depth"], f"depth_{typ}_reflect": t["depth_reflect"], f"opacity_{typ}": t["opacity"],
           f"x_surface_{typ}": t["x_surface"], "reflect_direction": t["reflect_direction"]}
    res[f"surface_normal_{typ}" if last.has_normal else f"surface_normal_grad_{typ}"] = t["surface_normal"]
    if last.has_mirror:
        res[f"mirror_mask_{typ}"] = t["mirror_mask"]
    if level_rays is not None:
        res["level_rays"] = level_rays
    if stats is not None:
        res["fused_stats"] = stats
    return res
