"""Host-side mirror of R/models/mirror_nerf.py: ``Embedding`` and ``MirrorNeRF`` with the reference's
constructor arguments, attribute names and ``state_dict`` keys (so reference checkpoints load unchanged,
R/utils/__init__.py:109-136), whose ``forward`` runs the CUDA field kernels of libmnrf.so.

``packed_field(module)`` turns any module that carries the reference parameter set (ours or the reference's own
class) into the device-resident packed weight object the kernels consume, cached on the module and re-packed
when a parameter's version counter or storage changes (optimizer step, ``load_state_dict``).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib

PARAM_KEYS = tuple(
    [f"xyz_encoding_{i + 1}.0.{wb}" for i in range(8) for wb in ("weight", "bias")]
    + ["xyz_encoding_final.weight", "xyz_encoding_final.bias", "dir_encoding.0.weight", "dir_encoding.0.bias",
       "sigma.weight", "sigma.bias", "rgb.0.weight", "rgb.0.bias",
       "normal_net.0.weight", "normal_net.0.bias", "normal_net.1.weight", "normal_net.1.bias",
       "is_mirror_net.0.weight", "is_mirror_net.0.bias", "is_mirror_net.2.weight", "is_mirror_net.2.bias"])
assert len(PARAM_KEYS) == _lib.NUM_PARAM_TENSORS

_EXPECTED_SHAPES = {
    "xyz_encoding_1.0.weight": (256, 63), "xyz_encoding_5.0.weight": (256, 319),
    "xyz_encoding_final.weight": (256, 256), "dir_encoding.0.weight": (128, 283), "sigma.weight": (1, 256),
    "rgb.0.weight": (3, 128), "normal_net.0.weight": (128, 256), "normal_net.1.weight": (3, 128),
    "is_mirror_net.0.weight": (128, 256), "is_mirror_net.2.weight": (1, 128)}
for _i in (2, 3, 4, 6, 7, 8):
    _EXPECTED_SHAPES[f"xyz_encoding_{_i}.0.weight"] = (256, 256)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class PackedField:
    """Owns one ``mnrf_field`` (device-resident packed weights of one MirrorNeRF)."""

    def __init__(self, tensors):
        self._lib = _lib.load()
        self.handle = C.c_void_p()
        arr = (C.c_void_p * _lib.NUM_PARAM_TENSORS)(*[None if t is None else t.data_ptr() for t in tensors])
        _lib.check(self._lib.mnrf_field_create(C.byref(self.handle), arr, _stream_ptr()), "mnrf_field_create")
        self.has_normal = bool(self._lib.mnrf_field_has_normal(self.handle))
        self.has_mirror = bool(self._lib.mnrf_field_has_mirror(self.handle))
        self.generation = 0  # bumped by every re-pack: autograd nodes refuse to run their backward against newer weights

    def update(self, tensors):
        arr = (C.c_void_p * _lib.NUM_PARAM_TENSORS)(*[None if t is None else t.data_ptr() for t in tensors])
        with torch.cuda.device(next(t for t in tensors if t is not None).device):
            _lib.check(self._lib.mnrf_field_update(self.handle, arr, _stream_ptr()), "mnrf_field_update")
        self.generation += 1

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self._lib.mnrf_field_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:  # interpreter shutdown
            pass


def _collect(source, with_key=False):
    """The 32 parameter tensors (or None for an absent head) of an nn.Module or a {key: tensor} mapping.  with_key: also the
    cache key, built from the ORIGINAL tensors' (data_ptr, _version) -- a `.contiguous()` temporary of a non-contiguous
    parameter can reuse an address with version 0 and would fake a cache hit."""
    named = dict(source.named_parameters()) if isinstance(source, nn.Module) else dict(source)
    out = []
    key = []
    for k in PARAM_KEYS:
        t = named.get(k)
        if t is None:
            if not (k.startswith("normal_net") or k.startswith("is_mirror_net")):
                raise KeyError(f"MirrorNeRF parameter '{k}' missing (only D=8, W=256, skips=[4] is supported)")
            out.append(None)
            key.append(None)
            continue
        if not t.is_cuda:
            raise RuntimeError(f"parameter '{k}' is on {t.device}: the renderer has no CPU path, move the model to CUDA")
        if t.dtype != torch.float32:
            raise RuntimeError(f"parameter '{k}' must be float32 (got {t.dtype})")
        exp = _EXPECTED_SHAPES.get(k)
        if exp is not None and tuple(t.shape) != exp:
            raise RuntimeError(f"parameter '{k}' has shape {tuple(t.shape)}, expected {exp} "
                               "(only the reference architecture D=8, W=256, skips=[4], 10/4 frequencies is supported)")
        out.append(t.detach() if t.is_contiguous() else t.detach().contiguous())
        key.append((t.data_ptr(), t._version, t.is_contiguous()))
    return (out, tuple(key)) if with_key else out


def packed_field(source) -> PackedField:
    """Packed device weights for ``source`` (module or state-dict-like mapping), cached on modules."""
    tensors, key = _collect(source, with_key=True)
    if isinstance(source, nn.Module):
        cached = source.__dict__.get("_mnrf_packed")
        if cached is not None:
            pf, old_key, dev = cached
            if dev == tensors[0].device:
                if old_key != key:
                    pf.update(tensors)
                    source.__dict__["_mnrf_packed"] = (pf, key, dev)
                return pf
    with torch.cuda.device(tensors[0].device):
        pf = PackedField(tensors)
    if isinstance(source, nn.Module):
        source.__dict__["_mnrf_packed"] = (pf, key, tensors[0].device)
    return pf


def invalidate_packed(module):
    """Force a re-pack at the next use (parameters were changed by a raw kernel that does not bump ``_version``)."""
    cached = module.__dict__.get("_mnrf_packed")
    if cached is not None:
        module.__dict__["_mnrf_packed"] = (cached[0], None, cached[2])


def _no_autograd(what, tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{what}: gradients are produced by render_rays (mirror_nerf_b200/autograd.py), not by this flat-batch "
            "forward; call it under torch.no_grad()")


class Embedding(nn.Module):
    """[x, sin(2^k x), cos(2^k x), ...] (R/models/mirror_nerf.py:6-38), computed by mnrf_embed on CUDA."""

    def __init__(self, N_freqs, logscale=True):
        super().__init__()
        if not logscale:
            raise NotImplementedError("only logscale=True frequency bands (the reference default) are supported")
        self.N_freqs = N_freqs
        self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs) if N_freqs > 0 else torch.zeros(0)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("Embedding.forward: input must be a CUDA tensor (no CPU path)")
        _no_autograd("Embedding.forward", [x])
        if x.shape[-1] != 3:
            raise RuntimeError("Embedding.forward: expected (..., 3) input")
        xf = x.reshape(-1, 3).contiguous().float()
        out = torch.empty(xf.shape[0], 3 + 6 * self.N_freqs, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            lib = _lib.load()
            _lib.check(lib.mnrf_embed(_ptr(xf), xf.shape[0], self.N_freqs, _ptr(out), _stream_ptr()), "mnrf_embed")
        return out.reshape(*x.shape[:-1], out.shape[-1])


class MirrorNeRF(nn.Module):
    """Same parameters / state_dict as R/models/mirror_nerf.py:41-99; ``forward`` follows :101-187."""

    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=[4], **kwargs):
        super().__init__()
        if (D, W, in_channels_xyz, in_channels_dir, list(skips)) != (8, 256, 63, 27, [4]):
            raise NotImplementedError("the B200 kernels implement the reference architecture only: "
                                      "D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=[4]")
        self.D, self.W = D, W
        self.in_channels_xyz, self.in_channels_dir, self.skips = in_channels_xyz, in_channels_dir, skips
        for i in range(D):
            fan_in = in_channels_xyz if i == 0 else (W + in_channels_xyz if i in skips else W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        self.geo_feat_dim = W
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2), nn.ReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        self.predict_normal = kwargs.get("predict_normal", False)
        if self.predict_normal:
            self.hidden_dim_normal = W // 2
            self.normal_net = nn.Sequential(nn.Linear(W, W // 2), nn.Linear(W // 2, 3))
        self.predict_mirror_mask = kwargs.get("predict_mirror_mask", False)
        if self.predict_mirror_mask:
            self.hidden_dim_is_mirror = W // 2
            self.is_mirror_net = nn.Sequential(nn.Linear(W, W // 2), nn.LeakyReLU(inplace=True),
                                               nn.Linear(W // 2, 1), nn.Sigmoid())
        self.field_impl = "tc3"  # kernel used when neither analytic normals nor geo_feat force the fp32 one
        self.return_geo_feat = True

    def forward(self, x, compute_normal=True, sigma_only=False, embedding_xyz=None, embedding_dir=None,
                mirror_mask=None, detach_density_outside_mirror_for_mask_loss=False,
                detach_density_for_mask_loss=False, detach_density_for_normal_loss=False):
        """x: (B, 3+27) = [xyz | embedded dir], or (B,3) when sigma_only.  Returns the reference's dict:
        sigma (B,1), geo_feat (B,256), normal? (B,3), pred_normal? (B,3), rgb? (B,3), is_mirror? (B,1).
        The detach_* flags only shape gradients and have no effect on forward values."""
        if not x.is_cuda:
            raise RuntimeError("MirrorNeRF.forward: input must be a CUDA tensor (no CPU path)")
        if embedding_xyz is not None and getattr(embedding_xyz, "N_freqs", 10) != 10:
            raise NotImplementedError("only N_emb_xyz=10 is supported")
        _no_autograd("MirrorNeRF.forward", [x] + list(self.parameters()))
        width = 3 if sigma_only else 3 + self.in_channels_dir
        if x.dim() != 2 or x.shape[1] != width:
            raise RuntimeError(f"MirrorNeRF.forward: expected x of shape (B,{width}), got {tuple(x.shape)}")
        x = x.detach().contiguous().float()
        B = x.shape[0]
        pf = packed_field(self)
        dev = x.device
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        sigma = new(B, 1)
        geo = new(B, 256) if self.return_geo_feat else None
        normal = new(B, 3) if compute_normal else None
        pred = new(B, 3) if pf.has_normal else None
        rgb = None if sigma_only else new(B, 3)
        mirror = None if (sigma_only or not pf.has_mirror) else new(B, 1)
        with torch.cuda.device(dev):
            lib = _lib.load()
            _lib.check(lib.mnrf_field_eval_points(pf.handle, _lib.IMPL_BY_NAME[self.field_impl], _ptr(x), B,
                                                  int(sigma_only), _ptr(sigma), _ptr(rgb), _ptr(mirror), _ptr(pred),
                                                  _ptr(normal), _ptr(geo), _stream_ptr()), "mnrf_field_eval_points")
        out = {}
        if compute_normal:
            out["normal"] = normal
        out["sigma"] = sigma
        if geo is not None:
            out["geo_feat"] = geo
        if pred is not None:
            out["pred_normal"] = pred
        if not sigma_only:
            out["rgb"] = rgb
            if mirror is not None:
                out["is_mirror"] = mirror
        return out
