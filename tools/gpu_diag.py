"""GPU diagnostics: prints error statistics of every CUDA stage against the CPU oracle / golden vectors.
Used during bring-up (run under gpurun); the asserting versions of these checks live in tests/.

    python tools/gpu_diag.py basic        # samplers, compositor, fp32 field kernel, render_rays(fp32)
    python tools/gpu_diag.py tc           # tcgen05 field kernel (tc3 + tc1) vs fp32 kernel and golden
    python tools/gpu_diag.py perf         # quick timing of the field kernels
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import T, err_stats, fmt_stats, make_models  # noqa: E402
from mirror_nerf_b200 import _lib  # noqa: E402
from mirror_nerf_b200.rendering import render_rays, sample_pdf  # noqa: E402
from mirror_nerf_b200.synthetic import make_state_dict, random_rays  # noqa: E402
from oracle import mirror_nerf_oracle as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return dict(np.load(os.path.join(G, name + ".npz")))


def stage(fn):
    print(f"\n=== {fn.__name__} ===", flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


def coarse_z():
    lib = _lib.load()
    rays = random_rays(1000, seed=5)
    for use_disp in (False, True):
        for perturb in (0.0, 1.0):
            u = torch.rand(1000, 64, generator=torch.Generator().manual_seed(3))
            want = O.coarse_z_vals(rays, 64, use_disp, perturb, u)
            from mirror_nerf_b200.rendering import _linspace
            r = rays.cuda()
            z = torch.empty(1000, 64, device="cuda")
            _lib.check(lib.mnrf_coarse_z(r.data_ptr(), 1000, _linspace(64, "cuda").data_ptr(), 64, int(use_disp), perturb,
                                         u.cuda().data_ptr(), z.data_ptr(), None))
            print(f"use_disp={use_disp} perturb={perturb}: bit-equal {torch.equal(z.cpu(), want)}  "
                  f"n_diff {(z.cpu() != want).sum().item()}  max {float((z.cpu() - want).abs().max()):.3e}")


def pdf():
    g = golden("sample_pdf")
    bins, w = T(g["bins"], "cuda"), T(g["weights"], "cuda")
    s, inds, cdf = sample_pdf(bins, w, 128, det=True, return_inds=True)
    print("cdf bit-equal", torch.equal(cdf.cpu(), T(g["cdf"])), " n_diff", (cdf.cpu() != T(g["cdf"])).sum().item())
    print("inds_det equal", torch.equal(inds.cpu(), T(g["inds_det"])))
    print("samples_det bit-equal", torch.equal(s.cpu(), T(g["det"])), " max", float((s.cpu() - T(g["det"])).abs().max()))
    lib = _lib.load()
    u = T(g["u"], "cuda").contiguous()
    inds2 = torch.empty(64, 128, device="cuda", dtype=torch.int64)
    c = T(g["cdf"], "cuda").contiguous()
    _lib.check(lib.mnrf_searchsorted_right(c.data_ptr(), 64, 63, u.data_ptr(), 128, 128, inds2.data_ptr(), None))
    print("searchsorted(rnd u) equal", torch.equal(inds2.cpu(), T(g["inds_rnd"])))


def field_fp32():
    g = golden("field")
    models, emb = make_models()
    m = models["fine"]
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
    with torch.no_grad():
        o = m(x, compute_normal=True, sigma_only=False)
    for k, gk in (("sigma", "full_sigma"), ("geo_feat", "full_geo_feat"), ("pred_normal", "full_pred_normal"),
                  ("rgb", "full_rgb"), ("is_mirror", "full_is_mirror"), ("normal", "grad_normal")):
        print(fmt_stats("fp32 " + k, err_stats(o[k].cpu(), T(g[gk]))))
    with torch.no_grad():
        o = m(T(g["xyz"]).cuda(), compute_normal=False, sigma_only=True)
    print(fmt_stats("fp32 sigonly sigma", err_stats(o["sigma"].cpu(), T(g["sigonly_sigma"]))))
    print(fmt_stats("fp32 sigonly pred_normal", err_stats(o["pred_normal"].cpu(), T(g["sigonly_pred_normal"]))))
    e = emb["xyz"](T(g["xyz"]).cuda())
    print("embed xyz max abs", float((e.cpu() - T(g["pe_xyz"])).abs().max()))


def render(impl):
    g = golden("render_eval")
    models, emb = make_models()
    with torch.no_grad():
        r = render_rays(models, emb, T(g["rays"]).cuda(), 64, False, 0, 0, 128, 32768, False, test_time=True,
                        compute_normal=False, field_impl=impl)
    print("keys equal", set(r) == set(g) - {"rays"}, sorted(set(r) ^ (set(g) - {"rays"})))
    for k in sorted(r):
        print(fmt_stats(f"{impl} {k}", err_stats(r[k].cpu(), T(g[k]))))
    print("z_vals_coarse bit-equal", torch.equal(r["z_vals_coarse"].cpu(), T(g["z_vals_coarse"])))


def render_fp32():
    render("fp32")


def render_train_fp32():
    g = golden("render_train")
    models, emb = make_models()
    rng = {k.split("/", 1)[1]: T(a) for k, a in g.items() if k.startswith("rng/")}
    with torch.no_grad():
        r = render_rays(models, emb, T(g["rays"]).cuda(), 64, False, 1.0, 1.0, 128, 32768, False, test_time=False,
                        compute_normal=True, rng=rng, field_impl="fp32")
    want = {k.split("/", 1)[1]: a for k, a in g.items() if k.startswith("out/")}
    print("keys equal", set(r) == set(want), sorted(set(r) ^ set(want)))
    for k in sorted(set(r) & set(want)):
        print(fmt_stats(f"train {k}", err_stats(r[k].cpu(), T(want[k]))))


def tc_field():
    g = golden("field")
    models, emb = make_models()
    m = models["fine"]
    m.return_geo_feat = False
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
    for impl in ("tc3", "tc1"):
        m.field_impl = impl
        with torch.no_grad():
            o = m(x, compute_normal=False, sigma_only=False)
        torch.cuda.synchronize()
        for k, gk in (("sigma", "full_sigma"), ("pred_normal", "full_pred_normal"), ("rgb", "full_rgb"),
                      ("is_mirror", "full_is_mirror")):
            print(fmt_stats(f"{impl} {k}", err_stats(o[k].cpu(), T(g[gk]))))
    print("sample sigma tc:", o["sigma"][:4, 0].tolist(), " ref:", g["full_sigma"][:4, 0].tolist())


def render_tc3():
    render("tc3")


def render_tc1():
    render("tc1")


def tc_big():
    """tc3 vs fp32 kernel on a larger batch with tail tiles and many tiles per CTA."""
    models, emb = make_models()
    rays = random_rays(3001, seed=9).cuda()
    with torch.no_grad():
        a = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False,
                        field_impl="fp32")
        b = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False,
                        field_impl="tc3")
    for k in sorted(a):
        print(fmt_stats(f"tc3~fp32 {k}", err_stats(b[k].cpu(), a[k].cpu())))


def perf():
    models, emb = make_models()
    n = 32768
    rays = random_rays(n, seed=4).cuda()
    for impl in ("tc3", "tc1"):
        with torch.no_grad():
            for _ in range(2):
                render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True,
                            compute_normal=False, field_impl=impl)
            torch.cuda.synchronize()
            t0 = time.time()
            reps = 5
            for _ in range(reps):
                render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True,
                            compute_normal=False, field_impl=impl)
            torch.cuda.synchronize()
            dt = (time.time() - t0) / reps
        print(f"{impl}: {n / dt:,.0f} rays/s  ({dt * 1e3:.2f} ms per {n} rays; "
              f"{n * 320.36e6 / dt / 1e12:.1f} algorithmic TFLOP/s)")


STAGES = {
    "basic": [coarse_z, pdf, field_fp32, render_fp32, render_train_fp32],
    "tc": [tc_field, render_tc3, render_tc1, tc_big],
    "perf": [perf],
}

if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0), " swap:", os.environ.get("MNRF_TC_DESC_SWAP"))
    for name in sys.argv[1:]:
        for fn in STAGES[name]:
            stage(fn)
