"""Gradient-enabled render_rays: what torch.autograd records for
R/models/rendering.py:54-369 when R/train.py:129-145 calls it with parameters that require grad.

One ``torch.autograd.Function`` per pass (field at the samples of a ray batch -> compositor): its forward runs
``mnrf_train_pass_fwd`` (all activations kept in a workspace tensor), its backward ``mnrf_train_pass_bwd`` (hand-written
backward, including the double backward through the analytic normals).  torch is only the plumbing: it owns the memory,
routes the output gradients into the Function and the parameter gradients into ``.grad``.

The same Function serves both field kinds: the MirrorNeRF MLP (32 parameter tensors, csrc/train.cu + train_tc.cu) and the
hash-grid field MirrorNeRFTcnn (12 parameter tensors incl. the hash table, csrc/train_hash.cu); the library dispatches on the
field handle.

Rays that require grad (secondary rays built from x_surface / surface normals, train.py:194-243) receive dL/d[o, d] as
well.  ``z_vals_*`` carry no gradient, as in the reference (rendering.py:335,353 detach the fine depths; the coarse depths
depend on near / far only).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from .mirror_nerf import PARAM_KEYS, _ptr, _stream_ptr, packed_field

# rays per autograd node: bounds the saved activations to ~ MAX_TRAIN_RAYS * (S) * 20 KB
MAX_TRAIN_RAYS = int(os.environ.get("MNRF_MAX_TRAIN_RAYS", 4096))

_OUT_ORDER = ("weights", "opacity", "rgb", "depth", "mirror_mask", "normal", "surface_normal_grad", "pred_normal",
              "surface_normal", "normal_dif", "x_surface")


def _present_params(module, all_keys=PARAM_KEYS):
    """(keys, tensors): the module's parameters in ABI order (absent heads skipped)."""
    named = dict(module.named_parameters())
    keys = [k for k in all_keys if k in named]
    return keys, [named[k] for k in keys]


class _PassFn(torch.autograd.Function):
    """outputs = pass(rays, z; params).  ``meta`` carries the non-tensor configuration."""

    @staticmethod
    def forward(ctx, meta, rays, z, noise, ray_detach, *params):
        lib = _lib.load()
        pf = meta["field"]
        n, S = z.shape
        dev = rays.device
        cn = int(meta["compute_normal"])
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        t = {"weights": new(n, S), "opacity": new(n), "rgb": new(n, 3), "depth": new(n), "x_surface": new(n, 3)}
        if pf.has_mirror:
            t["mirror_mask"] = new(n)
        if cn:
            t["normal"] = new(n, S, 3)
            t["surface_normal_grad"] = new(n, 3)
        if pf.has_normal:
            t["pred_normal"] = new(n, S, 3)
            t["surface_normal"] = new(n, 3)
            if cn:
                t["normal_dif"] = new(n)
        cfg = _lib.TrainCfg(S=S, compute_normal=cn, white_back=int(meta["white_back"]),
                            noise_std=float(meta["noise_std"]),
                            detach_density_for_mask_loss=int(meta["detach_mask"]),
                            detach_density_for_normal_loss=int(meta["detach_normal"]))
        with torch.cuda.device(dev):
            need = int(lib.mnrf_field_train_fwd_workspace_bytes(pf.handle, n, S, cn))
            ws = torch.empty(max(need, 1), device=dev, dtype=torch.uint8)
            p = lambda k: _ptr(t.get(k))
            out = _lib.CompositeOut(weights=p("weights"), opacity=p("opacity"), rgb=p("rgb"), depth=p("depth"),
                                    mirror_mask=p("mirror_mask"), pred_normal=p("pred_normal"),
                                    surface_normal=p("surface_normal"), surface_normal_grad=p("surface_normal_grad"),
                                    normal_dif=p("normal_dif"), x_surface=p("x_surface"))
            _lib.check(lib.mnrf_train_pass_fwd(pf.handle, _ptr(rays), _ptr(z), _ptr(noise), n, C.byref(cfg), _ptr(ws),
                                               need, C.byref(out), p("normal"), _stream_ptr()), "mnrf_train_pass_fwd")
        ctx.meta, ctx.cfg, ctx.ws, ctx.ws_bytes = meta, cfg, ws, need
        ctx.rays, ctx.z, ctx.noise, ctx.ray_detach = rays.detach(), z, noise, ray_detach
        ctx.depth = t["depth"].clone()  # private copy: callers may modify outputs in place
        ctx.param_shapes = [tuple(q.shape) for q in params]
        # The backward re-reads the packed weights through pf.handle.  save_for_backward makes autograd's version check cover
        # the parameters (an in-place optimizer step between forward and backward raises, as for any torch op), and the pack
        # generation covers re-packs by raw kernels that do not bump `_version` (FlatDataParallel.step, load_state_dict).
        ctx.save_for_backward(*params)
        ctx.generation = getattr(pf, "generation", 0)
        ctx.out_keys = [k for k in _OUT_ORDER if k in t]
        return tuple(t[k] for k in ctx.out_keys)

    @staticmethod
    @torch.autograd.function.once_differentiable  # the hand-written backward is not itself differentiable: fail loudly on create_graph
    def backward(ctx, *gouts):
        lib = _lib.load()
        meta, cfg = ctx.meta, ctx.cfg
        pf = meta["field"]
        _ = ctx.saved_tensors  # raises if a parameter was modified in place since the forward
        if getattr(pf, "generation", 0) != ctx.generation:
            raise RuntimeError("the field's weights were re-packed between this pass's forward and its backward (optimizer step / "
                               "load_state_dict in between): the saved activations no longer match the weights")
        n, S = ctx.z.shape
        dev = ctx.rays.device
        g = {}
        for k, go in zip(ctx.out_keys, gouts):
            if go is not None:
                g[k] = go.contiguous().float()
        grads = _lib.TrainGrads(**{k: _ptr(g.get(k)) for k in _lib.TRAIN_GRAD_FIELDS})
        # gradient tensors in the reference's parameter order / layout; absent heads stay NULL
        # one zero-filled flat buffer, one view per parameter (one fill launch instead of one per tensor)
        sizes = [int(torch.Size(shp).numel()) for shp in ctx.param_shapes]
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + ((sz + 3) & ~3))  # 16-byte aligned views (vector atomics / float4 stores of the kernels)
        flat = torch.zeros(offs[-1], device=dev, dtype=torch.float32)
        gts = {k: flat[o:o + sz].view(shp) for k, shp, o, sz in zip(meta["keys"], ctx.param_shapes, offs, sizes)}
        all_keys = meta["all_keys"]  # the ABI's tensor order for this field kind (32 MLP tensors / 12 hash-grid tensors)
        arr = (C.c_void_p * len(all_keys))(*[None if k not in gts else gts[k].data_ptr() for k in all_keys])
        grad_rays = torch.empty(n, 8, device=dev, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(dev):
            need = int(lib.mnrf_field_train_bwd_workspace_bytes(pf.handle, n, S, cfg.compute_normal))
            wsb = torch.empty(max(need, 1), device=dev, dtype=torch.uint8)
            _lib.check(lib.mnrf_train_pass_bwd(pf.handle, _ptr(ctx.rays), _ptr(ctx.z), _ptr(ctx.noise), n, C.byref(cfg),
                                               _ptr(ctx.ws), ctx.ws_bytes, _ptr(wsb), need, C.byref(grads),
                                               _ptr(ctx.ray_detach), arr, _ptr(ctx.depth), _ptr(grad_rays),
                                               _stream_ptr()), "mnrf_train_pass_bwd")
        return (None, grad_rays, None, None, None) + tuple(gts[k] for k in meta["keys"])


def run_pass(module, rays, z, noise, ray_detach, *, compute_normal, white_back, noise_std, detach_mask, detach_normal):
    """One differentiable pass.  Returns {name: tensor} (without the _coarse/_fine suffix)."""
    from .mirror_nerf_tcnn import HASH_PARAM_KEYS, is_hash_field, packed_hash_field
    hashed = is_hash_field(module)
    pf = packed_hash_field(module) if hashed else packed_field(module)
    all_keys = HASH_PARAM_KEYS if hashed else PARAM_KEYS
    keys, params = _present_params(module, all_keys)
    meta = dict(field=pf, keys=keys, all_keys=all_keys, compute_normal=compute_normal, white_back=white_back, noise_std=noise_std,
                detach_mask=detach_mask, detach_normal=detach_normal)
    outs = _PassFn.apply(meta, rays, z, noise, ray_detach, *params)
    names = [k for k in _OUT_ORDER if (k not in ("mirror_mask",) or pf.has_mirror)
             and (k not in ("normal", "surface_normal_grad") or compute_normal)
             and (k not in ("pred_normal", "surface_normal") or pf.has_normal)
             and (k != "normal_dif" or (pf.has_normal and compute_normal))]
    assert len(names) == len(outs)
    return dict(zip(names, outs))
