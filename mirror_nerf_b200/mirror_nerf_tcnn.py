"""Host-side mirror of R/models/mirror_nerf_tcnn.py (BASELINE config 3, `--model_type nerf_tcnn`): ``MirrorNeRFTcnn`` with
the reference's constructor arguments and ``state_dict`` keys (``encoder.params`` = tinycudann's flat hash-table parameter,
``sigma_net.{0,1}.weight``, ``color_net.{0,1,2}.weight``, ``normal_net.{0,1}.weight``, ``is_mirror_net.{0,2}.{weight,bias}``),
whose forward runs ``csrc/field_hash.cu``.  ``render_rays`` accepts these modules in ``models`` together with the identity
embeddings ``Embedding(0)`` the reference uses for this model (R/train.py:69-70).

The encoder follows tinycudann's HashGrid algorithm (parity unpinned, see oracle/hashgrid_oracle.py).  Gradients (hash table,
small MLPs, rays, double backward through the analytic normal) flow through ``render_rays`` (autograd.py -> csrc/train_hash.cu);
calling ``forward`` directly is inference only.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
from torch import nn

from . import _lib
from .mirror_nerf import _ptr, _stream_ptr

N_LEVELS, N_FEATURES, LOG2_HASHMAP, BASE_RES = 16, 2, 19, 16

HASH_PARAM_KEYS = ("encoder.params", "sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight",
                   "color_net.1.weight", "color_net.2.weight", "normal_net.0.weight", "normal_net.1.weight",
                   "is_mirror_net.0.weight", "is_mirror_net.0.bias", "is_mirror_net.2.weight", "is_mirror_net.2.bias")
_SHAPES = {"sigma_net.0.weight": (64, 32), "sigma_net.1.weight": (16, 64), "color_net.0.weight": (64, 31),
           "color_net.1.weight": (64, 64), "color_net.2.weight": (3, 64), "normal_net.0.weight": (64, 15),
           "normal_net.1.weight": (3, 64), "is_mirror_net.0.weight": (32, 15), "is_mirror_net.0.bias": (32,),
           "is_mirror_net.2.weight": (1, 32), "is_mirror_net.2.bias": (1,)}



def _libm_f32(name):
    """glibc's float32 `name` (what tinycudann's host code calls); numpy's float32 ufunc if libm cannot be loaded."""
    try:
        import ctypes
        f = getattr(ctypes.CDLL("libm.so.6"), name)
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
        return lambda x: np.float32(f(float(np.float32(x))))
    except Exception:
        return {"log2f": lambda x: np.log2(np.float32(x)), "exp2f": lambda x: np.exp2(np.float32(x))}[name]


_log2f, _exp2f = _libm_f32("log2f"), _libm_f32("exp2f")


def level_table(bound=1.0):
    """Per level (grid scale, resolution, first table entry, entries) and the total number of entries, as tinycudann builds
    them (grid.h): scale_l = 2^(l*log2(per_level_scale))*16 - 1 in fp32, resolution = ceil(scale)+1, entries = min(round_up(
    res^3, 8), 2^19); per_level_scale = 2^(log2(2048*bound/16)/15) (R/models/mirror_nerf_tcnn.py:38)."""
    # tinycudann does this arithmetic in float32 on the host (grid.h: per_level_scale is read into a float,
    # grid_scale = exp2f(level * log2f(per_level_scale)) * base_resolution - 1.0f); a double-precision log2 differs in the last
    # ulp of the scale and, for some bounds (e.g. 0.5), in the resolution of a level, i.e. in the indexing
    pls = np.float32(np.exp2(np.log2(2048 * bound / N_LEVELS) / (N_LEVELS - 1)))
    log2_pls = _log2f(pls)
    out, offset = [], 0
    for lvl in range(N_LEVELS):
        scale = float(_exp2f(np.float32(lvl) * log2_pls) * np.float32(BASE_RES) - np.float32(1.0))
        res = int(math.ceil(scale)) + 1
        n = min((res ** 3 + 7) // 8 * 8, 1 << LOG2_HASHMAP)
        out.append((scale, res, offset, n))
        offset += n
    return out, offset


def n_encoder_params(bound=1.0):
    return level_table(bound)[1] * N_FEATURES


class _HashGridParams(nn.Module):
    """Stands in for ``tcnn.Encoding``: owns the flat fp32 table under the key ``params``."""

    def __init__(self, n):
        super().__init__()
        self.params = nn.Parameter(torch.empty(n).uniform_(-1e-4, 1e-4))  # tinycudann's initialisation range


class PackedHashField:
    """Owns one hash-grid ``mnrf_field`` (device copy of the table + packed small MLPs)."""

    def __init__(self, tensors, bound):
        self._lib = _lib.load()
        self.handle = C.c_void_p()
        lv, total = level_table(bound)
        arr = (C.c_void_p * 12)(*[None if t is None else t.data_ptr() for t in tensors])
        scale = (C.c_float * N_LEVELS)(*[l[0] for l in lv])
        res = (C.c_int * N_LEVELS)(*[l[1] for l in lv])
        off = (C.c_uint32 * N_LEVELS)(*[l[2] for l in lv])
        size = (C.c_uint32 * N_LEVELS)(*[l[3] for l in lv])
        _lib.check(self._lib.mnrf_hash_field_create(C.byref(self.handle), arr, tensors[0].numel(), float(bound), scale, res,
                                                    off, size, _stream_ptr()), "mnrf_hash_field_create")
        self.has_normal = bool(self._lib.mnrf_field_has_normal(self.handle))
        self.has_mirror = bool(self._lib.mnrf_field_has_mirror(self.handle))
        self.kind = "hash"
        self.bound, self.table_floats = float(bound), int(tensors[0].numel())
        self.generation = 0  # bumped by every re-pack (autograd.py refuses a backward against newer weights)

    def update(self, tensors):
        """Re-pack in place after an optimizer step (same table size and head set)."""
        arr = (C.c_void_p * 12)(*[None if t is None else t.data_ptr() for t in tensors])
        _lib.check(self._lib.mnrf_field_update(self.handle, arr, _stream_ptr()), "mnrf_field_update")
        self.generation += 1

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self._lib.mnrf_field_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def is_hash_field(module) -> bool:
    return isinstance(module, nn.Module) and hasattr(module, "encoder") and hasattr(module, "sigma_net")


def _collect_hash(module):
    named = dict(module.named_parameters())
    out = []
    for k in HASH_PARAM_KEYS:
        t = named.get(k)
        if t is None:
            if not (k.startswith("normal_net") or k.startswith("is_mirror_net")):
                raise KeyError(f"MirrorNeRFTcnn parameter '{k}' missing")
            out.append(None)
            continue
        if not t.is_cuda:
            raise RuntimeError(f"parameter '{k}' is on {t.device}: the renderer has no CPU path, move the model to CUDA")
        if t.dtype != torch.float32:
            raise RuntimeError(f"parameter '{k}' must be float32 (got {t.dtype})")
        exp = _SHAPES.get(k)
        if exp is not None and tuple(t.shape) != exp:
            raise RuntimeError(f"parameter '{k}' has shape {tuple(t.shape)}, expected {exp} (only the reference's "
                               "nerf_tcnn configuration is supported: 2x64 sigma net, 3x64 colour net, geo_feat_dim 15)")
        out.append(t.detach().reshape(-1) if k == "encoder.params" else (t.detach() if t.is_contiguous() else t.detach().contiguous()))
    return out


def packed_hash_field(module) -> PackedHashField:
    """Packed device weights for a MirrorNeRFTcnn-shaped module, cached on it and rebuilt when a parameter changes."""
    tensors = _collect_hash(module)
    bound = float(getattr(module, "bound", 1.0))
    if tensors[0].numel() != n_encoder_params(bound):
        raise RuntimeError(f"encoder.params has {tensors[0].numel()} elements, expected {n_encoder_params(bound)} for "
                           f"bound={bound} (16 levels x 2 features, 2^19 hash map, base resolution 16)")
    key = tuple((None if t is None else (t.data_ptr(), t._version)) for t in tensors) + (bound,)
    cached = module.__dict__.get("_mnrf_packed")
    if cached is not None and cached[2] == tensors[0].device:
        if cached[1] == key:
            return cached[0]
        pf = cached[0]  # a None key = invalidated by a raw optimizer kernel (parallel.FlatDataParallel.step)
        if pf.bound == bound and pf.has_normal == (tensors[6] is not None) and pf.has_mirror == (tensors[8] is not None) \
                and pf.table_floats == tensors[0].numel():
            with torch.cuda.device(tensors[0].device):
                cached[0].update(tensors)  # parameters changed (optimizer step): re-pack in place
            module.__dict__["_mnrf_packed"] = (cached[0], key, tensors[0].device)
            return cached[0]
    with torch.cuda.device(tensors[0].device):
        pf = PackedHashField(tensors, bound)
    module.__dict__["_mnrf_packed"] = (pf, key, tensors[0].device)
    return pf


class MirrorNeRFTcnn(nn.Module):
    """Same parameters / state_dict as R/models/mirror_nerf_tcnn.py:13-149 for the configuration R/train.py:73-100 builds
    (hashgrid + sphere_harmonics, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, no background net)."""

    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", encoding_bg="hashgrid", num_layers=2,
                 hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, num_layers_bg=2, hidden_dim_bg=64,
                 bound=1, **kwargs):
        super().__init__()
        if (encoding, encoding_dir, num_layers, hidden_dim, geo_feat_dim, num_layers_color, hidden_dim_color) != \
                ("hashgrid", "sphere_harmonics", 2, 64, 15, 3, 64) or kwargs.get("bg_radius", 0):
            raise NotImplementedError("the B200 kernel implements the reference's nerf_tcnn configuration only "
                                      "(R/train.py:73-100): 2x64 sigma net, 3x64 colour net, geo_feat_dim 15, no background")
        self.bound = bound
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.encoder = _HashGridParams(n_encoder_params(bound))
        self.in_dim = N_LEVELS * N_FEATURES
        self.sigma_net = nn.ModuleList([nn.Linear(32, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.num_layers_color, self.hidden_dim_color, self.in_dim_dir = num_layers_color, hidden_dim_color, 16
        self.color_net = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False),
                                        nn.Linear(64, 3, bias=False)])
        self.bg_net = None
        self.predict_normal = kwargs.get("predict_normal", False)
        if self.predict_normal:
            self.num_layers_normal, self.hidden_dim_normal = num_layers, hidden_dim
            self.normal_net = nn.ModuleList([nn.Linear(15, 64, bias=False), nn.Linear(64, 3, bias=False)])
        self.predict_mirror_mask = kwargs.get("predict_mirror_mask", False)
        if self.predict_mirror_mask:
            self.hidden_dim_is_mirror = hidden_dim // 2
            self.is_mirror_net = nn.Sequential(nn.Linear(15, 32), nn.LeakyReLU(inplace=True), nn.Linear(32, 1), nn.Sigmoid())

    def forward(self, x, compute_normal=True, sigma_only=False, embedding_xyz=None, embedding_dir=None, mirror_mask=None,
                detach_density_outside_mirror_for_mask_loss=False, detach_density_for_mask_loss=False,
                detach_density_for_normal_loss=False):
        """x: (B, 3+3) = [xyz | d], or (B,3) when sigma_only.  Returns sigma (B,) [the reference's shape, mirror_nerf_tcnn.py:
        234], normal? (B,3, analytic), pred_normal? (B,3), rgb? (B,3), is_mirror? (B,1).  geo_feat is not exported."""
        if not x.is_cuda:
            raise RuntimeError("MirrorNeRFTcnn.forward: input must be a CUDA tensor (no CPU path)")
        if compute_normal and sigma_only:
            raise NotImplementedError("MirrorNeRFTcnn: analytic normals need the full pass (sigma_only=False)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("MirrorNeRFTcnn.forward is inference only (call it under torch.no_grad()); gradients of "
                                      "the hash-grid field flow through render_rays")
        width = 3 if sigma_only else 6
        if x.dim() != 2 or x.shape[1] != width:
            raise RuntimeError(f"MirrorNeRFTcnn.forward: expected x of shape (B,{width}), got {tuple(x.shape)}")
        x = x.detach().contiguous().float()
        B = x.shape[0]
        pf = packed_hash_field(self)
        new = lambda *s: torch.empty(*s, device=x.device, dtype=torch.float32)
        sigma = new(B)
        normal = new(B, 3) if compute_normal else None
        pred = new(B, 3) if pf.has_normal else None
        rgb = None if sigma_only else new(B, 3)
        mirror = None if (sigma_only or not pf.has_mirror) else new(B, 1)
        with torch.cuda.device(x.device):
            lib = _lib.load()
            _lib.check(lib.mnrf_field_eval_points(pf.handle, _lib.IMPL_FP32, _ptr(x), B, int(sigma_only), _ptr(sigma),
                                                  _ptr(rgb), _ptr(mirror), _ptr(pred), _ptr(normal), None, _stream_ptr()),
                       "mnrf_field_eval_points")
        out = {"sigma": sigma}
        if normal is not None:
            out["normal"] = normal
        if pred is not None:
            out["pred_normal"] = pred
        if not sigma_only:
            out["rgb"] = rgb
            if mirror is not None:
                out["is_mirror"] = mirror
        return out
