"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the hash-grid field of the reference (BASELINE config 3,
R/models/mirror_nerf_tcnn.py:13-259).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

PARITY UNPINNED for the encoder: the reference's multiresolution hash encoding is `tinycudann.Encoding` (R/models/
mirror_nerf_tcnn.py:10,39-49), a third-party CUDA extension that is neither vendored in /root/reference nor pinned to a
version (R/README.md:33 installs NVlabs/tiny-cuda-nn at HEAD) and cannot run in this container (no GPU, no wheel).  There is
no reference test, fixture or golden vector at this boundary.  `hashgrid_encode` restates tiny-cuda-nn's published algorithm
(include/tiny-cuda-nn/encodings/grid.h at v1.6/v1.7: `grid_scale`, `grid_resolution`, `pos_fract` with the 0.5 offset,
`grid_index` with dense indexing while the level fits and the coherent prime hash {1, 2654435761, 805459861} otherwise,
per-level parameter counts rounded up to 8 and capped at 2^log2_hashmap_size, linear interpolation, level-major output) in
fp32; tiny-cuda-nn itself evaluates the table and its output in fp16.  The spherical-harmonics direction encoding IS in the
reference tree (R/models/shencoder/src/shencoder.cu:51-78, degree 4) and is restated from there; the small bias-free MLPs are
plain nn.Linear layers in the reference (R/models/mirror_nerf_tcnn.py:52-149,218-259).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

N_LEVELS = 16
N_FEATURES = 2
LOG2_HASHMAP = 19
BASE_RES = 16
PRIMES = (1, 2654435761, 805459861)


def per_level_scale(bound=1.0):
    """R/models/mirror_nerf_tcnn.py:38."""
    return float(np.exp2(np.log2(2048 * bound / N_LEVELS) / (N_LEVELS - 1)))



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
    """[(scale, resolution, offset, size)] per level (grid.h: offset table construction) and the total entry count."""
    # float32 host arithmetic as in tinycudann's grid.h (per_level_scale read into a float, log2f, exp2f)
    pls = np.float32(per_level_scale(bound))
    log2_pls = _log2f(pls)
    out, offset = [], 0
    for lvl in range(N_LEVELS):
        scale = float(_exp2f(np.float32(lvl) * log2_pls) * np.float32(BASE_RES) - np.float32(1.0))
        res = int(math.ceil(scale)) + 1
        n = res ** 3
        n = (n + 7) // 8 * 8
        n = min(n, 1 << LOG2_HASHMAP)
        out.append((scale, res, offset, n))
        offset += n
    return out, offset


def n_encoder_params(bound=1.0):
    return level_table(bound)[1] * N_FEATURES


def hashgrid_encode(params, x01, bound=1.0):
    """params: flat (n_entries*2,) fp32 table; x01: (B,3) in [0,1].  Returns (B, 32) level-major features."""
    table = params.view(-1, N_FEATURES)
    levels, _ = level_table(bound)
    feats = []
    for scale, res, offset, size in levels:
        pos = x01 * scale + 0.5
        g = torch.floor(pos)
        frac = pos - g
        g = g.to(torch.int64)
        acc = torch.zeros(x01.shape[0], N_FEATURES, dtype=x01.dtype)
        for corner in range(8):
            w = torch.ones(x01.shape[0], dtype=x01.dtype)
            c = []
            for d in range(3):
                bit = (corner >> d) & 1
                w = w * (frac[:, d] if bit else (1 - frac[:, d]))
                c.append((g[:, d] + bit) & 0xFFFFFFFF)
            # grid_index: dense while the stride fits into the level, hash otherwise
            stride, index, d = 1, torch.zeros_like(c[0]), 0
            while d < 3 and stride <= size:
                index = (index + c[d] * stride) & 0xFFFFFFFF
                stride *= res
                d += 1
            if size < stride:
                index = torch.zeros_like(c[0])
                for dd in range(3):
                    index = index ^ ((c[dd] * PRIMES[dd]) & 0xFFFFFFFF)
            index = index % size
            acc = acc + w.unsqueeze(-1) * table[offset + index]
        feats.append(acc)
    return torch.cat(feats, -1)


def sh4(d):
    """Real spherical harmonics up to degree 4 = 16 coefficients (R/models/shencoder/src/shencoder.cu:51-78)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814),
         -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
         0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
         0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
         0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, -1)


def l2_normalize(x):
    eps = torch.as_tensor(torch.finfo(torch.float32).eps)
    return x / torch.sqrt(torch.maximum(torch.sum(x ** 2, -1, keepdim=True), eps))


def param_shapes(bound=1.0, predict_normal=True, predict_mirror_mask=True):
    """state_dict keys / shapes of R/models/mirror_nerf_tcnn.py with the defaults train.py:73-100 passes."""
    s = OrderedDict()
    s["encoder.params"] = (n_encoder_params(bound),)
    s["sigma_net.0.weight"] = (64, 32)
    s["sigma_net.1.weight"] = (16, 64)
    s["color_net.0.weight"] = (64, 31)
    s["color_net.1.weight"] = (64, 64)
    s["color_net.2.weight"] = (3, 64)
    if predict_normal:
        s["normal_net.0.weight"] = (64, 15)
        s["normal_net.1.weight"] = (3, 64)
    if predict_mirror_mask:
        s["is_mirror_net.0.weight"] = (32, 15)
        s["is_mirror_net.0.bias"] = (32,)
        s["is_mirror_net.2.weight"] = (1, 32)
        s["is_mirror_net.2.bias"] = (1,)
    return s


def field_forward(p, x, bound=1.0, sigma_only=False, compute_normal=False, mirror_mask=None,
                  detach_density_outside_mirror_for_mask_loss=False, detach_density_for_mask_loss=False,
                  detach_density_for_normal_loss=False):
    """MirrorNeRFTcnn.forward (mirror_nerf_tcnn.py:151-259).  x: (B,6) = [xyz | d] or (B,3).  compute_normal: the analytic
    normal normalize(-d sigma/d xyz) by autograd through the restated encoder (mirror_nerf_tcnn.py:170-178); when gradients are
    being recorded (a parameter or x requires grad) the graph of that derivative is kept (create_graph=True, R/utils/func.py:
    10-25), so losses on the normal reach the table and sigma_net by double backward, as in the reference."""
    xyz = x[:, :3]
    out = {}
    train = torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in p.values()))
    if compute_normal:
        with torch.enable_grad():
            if not (train and xyz.requires_grad):
                xyz = xyz.detach().clone().requires_grad_(True)
            h = hashgrid_encode(p["encoder.params"], (xyz + bound) / (2 * bound), bound)
            h = F.linear(F.relu(F.linear(h, p["sigma_net.0.weight"])), p["sigma_net.1.weight"])
            sig = h[:, 0:1]
            g = torch.autograd.grad(sig, xyz, torch.ones_like(sig), retain_graph=True, create_graph=train)[0]
        if train:
            out["normal"] = l2_normalize(-g)
        else:
            out["normal"] = l2_normalize(-g.detach())
            h = h.detach()
    else:
        h = hashgrid_encode(p["encoder.params"], (xyz + bound) / (2 * bound), bound)
        h = F.linear(F.relu(F.linear(h, p["sigma_net.0.weight"])), p["sigma_net.1.weight"])
    out.update({"sigma": h[:, 0:1], "geo_feat": h[:, 1:]})  # sigma raw (mirror_nerf_tcnn.py:233-234), shaped (B,1) like the MLP field
    geo = out["geo_feat"]
    if "normal_net.0.weight" in p:
        gn = geo.detach() if detach_density_for_normal_loss else geo  # mirror_nerf_tcnn.py:186-191
        nh = F.linear(F.relu(F.linear(gn, p["normal_net.0.weight"])), p["normal_net.1.weight"])
        out["pred_normal"] = l2_normalize(nh)
    if not sigma_only:
        c = torch.cat([sh4(x[:, 3:6]), geo], -1)
        c = F.relu(F.linear(c, p["color_net.0.weight"]))
        c = F.relu(F.linear(c, p["color_net.1.weight"]))
        out["rgb"] = torch.sigmoid(F.linear(c, p["color_net.2.weight"]))
        if "is_mirror_net.0.weight" in p:
            gm = geo
            if detach_density_for_mask_loss:  # mirror_nerf_tcnn.py:199-216
                gm = geo.detach()
            elif (detach_density_outside_mirror_for_mask_loss and mirror_mask is not None
                  and not bool((mirror_mask < 0).any())):
                keep = mirror_mask.clone().bool().unsqueeze(-1)
                gm = torch.where(keep, geo, geo.detach())
            m = F.leaky_relu(F.linear(gm, p["is_mirror_net.0.weight"], p["is_mirror_net.0.bias"]), 0.01)
            out["is_mirror"] = torch.sigmoid(F.linear(m, p["is_mirror_net.2.weight"], p["is_mirror_net.2.bias"]))
    return out


def make_state_dict(seed=0, bound=1.0, sigma_scale=20.0, predict_normal=True, predict_mirror_mask=True, table_scale=1.0):
    """Synthetic weights: table ~ U(-1,1)*table_scale (tiny-cuda-nn initialises U(-1e-4,1e-4); a trained table is O(1)),
    nn.Linear-style layers; the sigma row of sigma_net.1 is scaled so that rays saturate."""
    g = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for k, shp in param_shapes(bound, predict_normal, predict_mirror_mask).items():
        if k == "encoder.params":
            sd[k] = torch.from_numpy((g.uniform(-1.0, 1.0, size=shp) * table_scale).astype(np.float32))
        else:
            fan_in = shp[-1] if k.endswith("weight") else param_shapes(bound)[k[:-4] + "weight"][1]
            b = 1.0 / math.sqrt(fan_in)
            sd[k] = torch.from_numpy(g.uniform(-b, b, size=shp).astype(np.float32))
    sd["sigma_net.1.weight"][0] *= float(sigma_scale)
    return sd
