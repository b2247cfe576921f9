"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the reference-shaped Python
API -> ctypes -> C ABI of libmnrf.so, against (a) golden vectors produced by the unmodified reference and (b) the CPU
oracle on the same seeded inputs.

Tolerances (north-star: ray/sample indices bit-exact; rgb/depth within 1e-3 relative fp32):
  * integer / index / depth-sampling work (coarse z, cdf, searchsorted indices, resampled depths given the same
    weights, compaction order): BIT-EXACT;
  * fp32 field outputs per point: relative error vs the tensor's RMS scale, median <= 1e-5 and p99 <= 1e-3
    (fp32 kernel: p99 <= 1e-4);
  * per-ray composited outputs on the adversarial golden field (sigma head x40, SURVEY.md 7.3: the reference itself is
    discontinuous in sign(sigma_last) and re-ordering fp32 sums already moves 0.05-0.15 % of rays by > 1e-3):
    median <= 2e-5 and at most 1 % of entries off by more than 1e-3 (relative to max(|ref|, rms)).

Every bound below is "measured x 2" (round 2: the measured distributions of a B200 run are in profiles/r02_*_parity_stats.json;
tests/util.py::err_stats logs them on every run into gpurun_out/parity_stats.json).  Where a fixture has only 8 or 12 rays the
fraction bound is "one ray": the FP32 CUDA-core kernel shows the same single ray (error 1.1e-3 .. 2e-3) on those fixtures.
"""
import pytest
import torch

from util import T, err_stats, fmt_stats, make_models

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm():
    return make_models()


@pytest.fixture(scope="module")
def oracle():
    from oracle import mirror_nerf_oracle as O
    return O


@pytest.fixture(scope="module")
def params():
    from mirror_nerf_b200.synthetic import scene_state_dicts
    return scene_state_dicts()


def assert_close_dist(got, want, name, median=2e-5, frac=0.01, p99=None):
    s = err_stats(got, want, name=name)
    msg = fmt_stats(name, s)
    assert s["median"] <= median, msg
    assert s["frac"] <= frac, msg
    if p99 is not None:
        assert s["p99"] <= p99, msg


# ---------------------------------------------------------------------------------------------- samplers
@pytest.mark.parametrize("use_disp", [False, True])
@pytest.mark.parametrize("perturb", [0.0, 1.0, 0.5])
def test_coarse_z_bit_exact(oracle, use_disp, perturb):
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.rendering import _linspace
    from mirror_nerf_b200.synthetic import random_rays
    lib = _lib.load()
    n, S = 777, 64
    rays = random_rays(n, seed=5)
    u = torch.rand(n, S, generator=torch.Generator().manual_seed(3))
    want = oracle.coarse_z_vals(rays, S, use_disp, perturb, u)
    z = torch.empty(n, S, device="cuda")
    r, uc = rays.cuda(), u.cuda()
    _lib.check(lib.mnrf_coarse_z(r.data_ptr(), n, _linspace(S, "cuda").data_ptr(), S, int(use_disp), perturb,
                                 uc.data_ptr(), z.data_ptr(), None))
    assert torch.equal(z.cpu(), want)


def test_embedding(golden, mm):
    g = golden("field")
    _, emb = mm
    for name, key in (("xyz", "xyz"), ("dir", "dir")):
        e = emb[name](T(g[key], "cuda")).cpu()
        want = T(g["pe_" + key])
        assert e.shape == want.shape
        # sin/cos of arguments up to 2^9*8: both libraries are within a few ulp of the exact value
        assert float((e - want).abs().max()) <= 1e-6


def test_searchsorted_bit_exact(golden):
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    g = golden("sample_pdf")
    cdf = T(g["cdf"], "cuda").contiguous()
    n, m = cdf.shape
    for ukey, ikey in (("u", "inds_rnd"),):
        u = T(g[ukey], "cuda").contiguous()
        inds = torch.empty(n, u.shape[1], device="cuda", dtype=torch.int64)
        _lib.check(lib.mnrf_searchsorted_right(cdf.data_ptr(), n, m, u.data_ptr(), u.shape[1], u.shape[1],
                                               inds.data_ptr(), None))
        assert torch.equal(inds.cpu(), T(g[ikey]))
    u = torch.linspace(0, 1, 128, device="cuda")
    inds = torch.empty(n, 128, device="cuda", dtype=torch.int64)
    _lib.check(lib.mnrf_searchsorted_right(cdf.data_ptr(), n, m, u.data_ptr(), 128, 0, inds.data_ptr(), None))
    assert torch.equal(inds.cpu(), T(g["inds_det"]))


def test_sample_pdf_golden_bit_exact(golden):
    """cdf, bin indices and resampled depths are bit-identical to the reference given the same weights (includes the
    all-zero row, single spike, two spikes and 1e-7 rows of the fixture)."""
    from mirror_nerf_b200.rendering import sample_pdf
    g = golden("sample_pdf")
    s, inds, cdf = sample_pdf(T(g["bins"], "cuda"), T(g["weights"], "cuda"), 128, det=True, return_inds=True)
    assert torch.equal(cdf.cpu(), T(g["cdf"]))
    assert torch.equal(inds.cpu(), T(g["inds_det"]))
    assert torch.equal(s.cpu(), T(g["det"]))


def test_sample_pdf_random_u_and_merge(oracle):
    """Fused resample + sort-merge kernel with per-ray random u against the oracle (bit-exact), ragged sizes."""
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(21)
    for n, S, ni in ((37, 64, 128), (5, 32, 16), (1, 64, 64), (130, 17, 5)):
        z = torch.sort(torch.rand(n, S, generator=gen) * 7 + 0.05, dim=1)[0]
        w = torch.rand(n, S, generator=gen) ** 6
        w[0] = 0
        u = torch.rand(n, ni, generator=gen)
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        smp, inds, cdf = oracle.sample_pdf(mid, w[:, 1:-1], ni, u=u, return_inds=True)
        want = torch.sort(torch.cat([z, smp], -1), -1)[0]
        zc, wc, uc = z.cuda(), w.cuda(), u.cuda()
        zf = torch.empty(n, S + ni, device="cuda")
        so = torch.empty(n, ni, device="cuda")
        io = torch.empty(n, ni, device="cuda", dtype=torch.int64)
        _lib.check(lib.mnrf_sample_pdf(zc.data_ptr(), wc.data_ptr(), n, S, ni, uc.data_ptr(), ni, zf.data_ptr(),
                                       so.data_ptr(), io.data_ptr(), None, None))
        assert torch.equal(io.cpu(), inds), (n, S, ni)
        assert torch.equal(so.cpu(), smp), (n, S, ni)
        assert torch.equal(zf.cpu(), want), (n, S, ni)
        assert bool((zf[:, 1:] >= zf[:, :-1]).all())


# ---------------------------------------------------------------------------------------------- field
def test_field_fp32_kernel_golden(golden, mm):
    g = golden("field")
    models, _ = mm
    m = models["fine"]
    m.return_geo_feat = True
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
    with torch.no_grad():
        o = m(x, compute_normal=True, sigma_only=False)
    assert list(o) == ["normal", "sigma", "geo_feat", "pred_normal", "rgb", "is_mirror"]
    for k, gk in (("sigma", "full_sigma"), ("geo_feat", "full_geo_feat"), ("pred_normal", "full_pred_normal"),
                  ("rgb", "full_rgb"), ("is_mirror", "full_is_mirror")):
        assert o[k].shape == T(g[gk]).shape
        assert_close_dist(o[k].cpu(), T(g[gk]), "fp32 " + k, median=2e-6, frac=0.0, p99=1e-4)
    # analytic normal: explicit reverse chain vs autograd (unit vectors: compare directions)
    cos = (o["normal"].cpu() * T(g["grad_normal"])).sum(-1)
    assert float(cos.min()) > 1 - 1e-4
    with torch.no_grad():
        o = m(T(g["xyz"]).cuda(), compute_normal=False, sigma_only=True)
    assert set(o) == {"sigma", "geo_feat", "pred_normal"}
    assert_close_dist(o["sigma"].cpu(), T(g["sigonly_sigma"]), "fp32 sigonly", median=2e-6, frac=0.0, p99=1e-4)
    assert_close_dist(o["pred_normal"].cpu(), T(g["sigonly_pred_normal"]), "fp32 sigonly pn", median=2e-6, frac=0.0,
                      p99=1e-4)


@pytest.mark.parametrize("impl,med,p99", [("tc3", 1e-5, 1e-3), ("tc1", 2e-3, 5e-2)])
def test_field_tcgen05_kernel_golden(golden, mm, impl, med, p99):
    g = golden("field")
    models, _ = mm
    m = models["fine"]
    m.return_geo_feat = False
    m.field_impl = impl
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
    try:
        with torch.no_grad():
            o = m(x, compute_normal=False, sigma_only=False)
    finally:
        m.field_impl = "tc3"
        m.return_geo_feat = True
    for k, gk in (("sigma", "full_sigma"), ("pred_normal", "full_pred_normal"), ("rgb", "full_rgb"),
                  ("is_mirror", "full_is_mirror")):
        assert_close_dist(o[k].cpu(), T(g[gk]), f"{impl} {k}", median=med, frac=1.0, p99=p99)


@pytest.mark.parametrize("impl", ["tc3", "tc1"])
def test_field_analytic_normal_on_tensor_cores(golden, mm, impl):
    """normalize(-d sigma/d xyz) (mirror_nerf.py:136-146) from the tcgen05 reverse chain vs the reference's autograd."""
    g = golden("field")
    models, _ = mm
    m = models["fine"]
    m.return_geo_feat = False
    m.field_impl = impl
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
    try:
        with torch.no_grad():
            o = m(x, compute_normal=True, sigma_only=False)
    finally:
        m.field_impl = "tc3"
        m.return_geo_feat = True
    cos = (o["normal"].cpu() * T(g["grad_normal"])).sum(-1)
    # d sigma/d xyz is discontinuous wherever a pre-activation crosses zero: a unit whose pre-activation is within rounding
    # of 0 flips its relu' bit and moves the normal by a finite angle (1 of the 192 golden points does for tc3: cos 0.957,
    # same point for tc1), so the check is on the distribution, not on the minimum
    if impl == "tc3":
        assert float((cos < 1 - 1e-4).float().mean()) <= 0.02, cos.min()
        assert float(cos.median()) > 1 - 1e-6
        assert_close_dist(o["normal"].cpu(), T(g["grad_normal"]), "tc3 analytic normal", median=1e-5, frac=0.02)
    else:
        assert float((cos < 0.99).float().mean()) <= 0.02 and float(cos.median()) > 1 - 1e-4, cos.min()
    if impl == "tc3":
        assert_close_dist(o["sigma"].cpu(), T(g["grad_sigma"]), "tc3 sigma (normals build)", median=1e-5, frac=0.0)
        assert_close_dist(o["rgb"].cpu(), T(g["grad_rgb"]), "tc3 rgb (normals build)", median=1e-5, frac=0.0)


def test_field_heads_optional(oracle):
    """MirrorNeRF default (no normal / mirror heads): keys and values."""
    from mirror_nerf_b200.synthetic import random_rays, scene_state_dicts
    models, emb = make_models(predict_normal=False, predict_mirror_mask=False)
    p = scene_state_dicts(predict_normal=False, predict_mirror_mask=False)
    rays = random_rays(40, seed=8)
    from mirror_nerf_b200.rendering import render_rays
    with torch.no_grad():
        got = render_rays(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False, test_time=True,
                          compute_normal=False)
        want = oracle.render_rays(p, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
    assert set(got) == set(want)
    assert_close_dist(got["rgb_fine"].cpu(), want["rgb_fine"], "noheads rgb")
    assert_close_dist(got["depth_fine"].cpu(), want["depth_fine"], "noheads depth")


# ---------------------------------------------------------------------------------------------- compositor
@pytest.mark.parametrize("S,white_back,noise_std", [(64, False, 0.0), (192, True, 0.0), (192, False, 1.0), (33, False, 0.5)])
def test_composite_vs_oracle(oracle, S, white_back, noise_std):
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(S)
    n = 101
    z = torch.sort(torch.rand(n, S, generator=gen) * 7 + 0.05, dim=1)[0]
    sig = torch.randn(n, S, generator=gen) * 20
    rgb = torch.rand(n, S, 3, generator=gen)
    mir = torch.rand(n, S, generator=gen)
    pn = torch.nn.functional.normalize(torch.randn(n, S, 3, generator=gen), dim=-1)
    nr = torch.nn.functional.normalize(torch.randn(n, S, 3, generator=gen), dim=-1)
    noise = torch.randn(n, S, generator=gen)
    rays = torch.randn(n, 8, generator=gen)
    want = {}
    oracle.composite(want, "x", z, sig, rgb, mir, nr, pn, noise_std=noise_std, white_back=white_back,
                     weights_only=False, noise=noise)
    raw = torch.cat([sig[..., None], rgb, mir[..., None], pn], -1).contiguous().cuda()
    d = lambda *s: torch.empty(*s, device="cuda")
    o = dict(weights=d(n, S), opacity=d(n), rgb=d(n, 3), depth=d(n), mirror_mask=d(n), pred_normal=d(n, S, 3),
             surface_normal=d(n, 3), surface_normal_grad=d(n, 3), normal_dif=d(n), x_surface=d(n, 3))
    st = _lib.CompositeOut(**{k: v.data_ptr() for k, v in o.items()})
    zc, nc, noc, rc = z.cuda(), nr.contiguous().cuda(), noise.cuda(), rays.cuda()
    import ctypes as C
    _lib.check(lib.mnrf_composite(rc.data_ptr(), zc.data_ptr(), raw.data_ptr(), 8, raw.data_ptr(), nc.data_ptr(),
                                  noc.data_ptr(), noise_std, n, S, int(white_back), C.byref(st), None))
    for k, wk in (("weights", "weights_x"), ("opacity", "opacity_x"), ("rgb", "rgb_x"), ("depth", "depth_x"),
                  ("mirror_mask", "mirror_mask_x"), ("surface_normal", "surface_normal_x"),
                  ("surface_normal_grad", "surface_normal_grad_x"), ("normal_dif", "normal_dif_x")):
        assert_close_dist(o[k].cpu(), want[wk], f"composite {k}", median=1e-6, frac=0.0, p99=1e-5)
    assert torch.equal(o["pred_normal"].cpu(), pn)
    xs = rays[:, :3] + rays[:, 3:6] * want["depth_x"][:, None]
    assert_close_dist(o["x_surface"].cpu(), xs, "x_surface", median=1e-6, frac=0.0, p99=1e-5)


# ---------------------------------------------------------------------------------------------- render_rays
@pytest.mark.parametrize("impl", ["fp32", "tc3"])
def test_render_eval_golden(golden, mm, impl):
    """BASELINE config-2 per-level call: 64 + 128 samples, eval mode, both heads."""
    from mirror_nerf_b200.rendering import render_rays
    g = golden("render_eval")
    models, emb = mm
    with torch.no_grad():
        r = render_rays(models, emb, T(g["rays"], "cuda"), 64, False, 0, 0, 128, 32768, False, test_time=True,
                        compute_normal=False, field_impl=impl)
    assert set(r) == set(g) - {"rays"}
    assert torch.equal(r["z_vals_coarse"].cpu(), T(g["z_vals_coarse"]))
    for k in sorted(r):
        assert tuple(r[k].shape) == g[k].shape, k
        assert r[k].dtype == torch.float32
        assert_close_dist(r[k].cpu(), T(g[k]), f"{impl} {k}")


VARIANTS = {
    "white_back": dict(args=(64, False, 0, 0, 128, 32768, True), kw=dict(test_time=True), fine=True),
    "use_disp": dict(args=(64, True, 0, 0, 128, 32768, False), kw=dict(test_time=True), fine=True),
    "coarse_only": dict(args=(64, False, 0, 0, 0, 32768, False), kw=dict(test_time=True), fine=False),
    "one_field": dict(args=(64, False, 0, 0, 128, 32768, False),
                      kw=dict(test_time=True, only_one_field=True, current_epoch=3), fine=False),
    "train_nonormal": dict(args=(64, False, 0, 0, 128, 32768, False), kw=dict(test_time=False), fine=True),
    "s32_i16": dict(args=(32, False, 0, 0, 16, 1000, False), kw=dict(test_time=True), fine=True),
}


@pytest.mark.parametrize("impl", ["fp32", "tc3"])
@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_render_variants_golden(golden, mm, tag, impl):
    from mirror_nerf_b200.rendering import render_rays
    g = golden("render_variants")
    v = VARIANTS[tag]
    models, emb = mm
    ms = models if v["fine"] else {"coarse": models["coarse"]}
    with torch.no_grad():
        r = render_rays(ms, emb, T(g["rays"], "cuda"), *v["args"], compute_normal=False, field_impl=impl, **v["kw"])
    want = {k.split("/", 1)[1]: a for k, a in g.items() if k.startswith(tag + "/")}
    assert set(r) == set(want), sorted(set(r) ^ set(want))
    for k in sorted(r):
        assert tuple(r[k].shape) == want[k].shape, k
        # 12 rays only: one ray (x channels) sits on the sigma_last discontinuity -- also for the fp32 kernel (measured: that ray is
        # off by 1.1e-3 (fp32) / 1.5e-3 (tc3) in opacity, every other entry < 1e-3; medians <= 1.4e-5)
        assert_close_dist(r[k].cpu(), T(want[k]), f"{tag}/{impl} {k}", median=3e-5, frac=0.09)


@pytest.mark.parametrize("impl", ["fp32", "tc3"])
def test_render_train_mode_forward_golden(golden, mm, impl):
    """test_time=False, compute_normal=True, perturb=1, noise_std=1 with the reference's RNG draws replayed
    (analytic normals: explicit fp32 chain / tcgen05 chain)."""
    from mirror_nerf_b200.rendering import render_rays
    g = golden("render_train")
    models, emb = mm
    rng = {k.split("/", 1)[1]: T(a) for k, a in g.items() if k.startswith("rng/")}
    with torch.no_grad():
        r = render_rays(models, emb, T(g["rays"], "cuda"), 64, False, 1.0, 1.0, 128, 32768, False, test_time=False,
                        compute_normal=True, rng=rng, field_impl=impl)
    want = {k.split("/", 1)[1]: a for k, a in g.items() if k.startswith("out/")}
    assert set(r) == set(want), sorted(set(r) ^ set(want))
    assert torch.equal(r["z_vals_coarse"].cpu(), T(want["z_vals_coarse"]))
    for k in sorted(r):
        assert tuple(r[k].shape) == want[k].shape, k
        # 8 rays, analytic normals of the sharp field, sigma noise: one or two rays (12.5 / 20.8 % of a 24-entry tensor) beyond 1e-3
        # depending on the summation order of the sigma head (a last-sample alpha flips with sign(sigma + noise), the
        # reference's own discontinuity); medians measured <= 9.8e-5.  The kernel is deterministic run to run
        # (test_gpu_round2.py::test_analytic_normal_kernel_is_deterministic).
        # The tight train-mode check is test_room_train_mode_forward below (scene-like field, 256 rays).
        assert_close_dist(r[k].cpu(), T(want[k]), f"train {k}", median=2e-4, frac=0.25)


def test_single_ray_and_empty(mm):
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    models, emb = mm
    with torch.no_grad():
        r1 = render_rays(models, emb, random_rays(1, seed=3).cuda(), 64, False, 0, 0, 128, 32768, False, test_time=True,
                         compute_normal=False)
        r0 = render_rays(models, emb, torch.zeros(0, 8, device="cuda"), 64, False, 0, 0, 128, 32768, False,
                         test_time=True, compute_normal=False)
    assert r1["rgb_fine"].shape == (1, 3) and r1["x_surface_fine"].shape == (1, 3)
    assert bool(torch.isfinite(r1["rgb_fine"]).all())
    assert r0["rgb_fine"].shape == (0, 3) and r0["weights_fine"].shape == (0, 192)


def test_full_size_properties(mm):
    """Size-independent properties on a batch big enough for several tiles per SM with a ragged tail:
    tc3 agrees with the fp32 kernel, weights sum to opacity, depths sorted, outputs independent of batch split."""
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import camera_rays
    models, emb = mm
    rays = camera_rays(800, 800)[::19][:30011].contiguous().cuda()
    with torch.no_grad():
        a = render_rays(models, emb, rays, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
        b = render_rays(models, emb, rays[:4099], 64, False, 0, 0, 128, 32768, False, test_time=True,
                        compute_normal=False, field_impl="fp32")
        c = render_rays(models, emb, rays[1000:2000], 64, False, 0, 0, 128, 32768, False, test_time=True,
                        compute_normal=False)
    for k in a:
        assert bool(torch.isfinite(a[k]).all()), k
    assert bool((a["z_vals_fine"][:, 1:] >= a["z_vals_fine"][:, :-1]).all())
    assert_close_dist(a["weights_fine"].sum(1).cpu(), a["opacity_fine"].cpu(), "sum w", median=1e-6, frac=0.0)
    assert bool((a["opacity_fine"] <= 1 + 1e-4).all())
    for k in ("rgb_fine", "depth_fine", "mirror_mask_fine", "surface_normal_fine", "opacity_fine"):
        assert_close_dist(a[k][:4099].cpu(), b[k].cpu(), f"tc3~fp32 {k}")
    for k in a:  # a batch is a set of independent rays: results do not depend on how it is split
        assert torch.equal(a[k][1000:2000], c[k]), k


# ---------------------------------------------------------------------------------------------- bounce
def test_reflect_compact_blend_vs_oracle(oracle):
    from mirror_nerf_b200.trace import blend_reflection, compact_rows, reflect_rays
    gen = torch.Generator().manual_seed(2)
    n = 5000
    rays = torch.randn(n, 8, generator=gen)
    xs = torch.randn(n, 3, generator=gen)
    nrm = torch.randn(n, 3, generator=gen)
    mask = torch.rand(n, generator=gen)
    mask[7] = 0.5
    want_sec, want_r = oracle.reflect_rays(rays, xs, nrm)
    m = mask.clone().cuda()
    sec, refl, flag = reflect_rays(rays.cuda(), xs.cuda(), nrm.cuda(), m)
    wm = mask.clone()
    wm[wm > 0.5] = 1
    wm[wm < 0.5] = 0
    assert torch.equal(m.cpu(), wm) and int(flag.item()) == 1
    assert_close_dist(sec.cpu(), want_sec, "secondary", median=1e-6, frac=0.0, p99=1e-5)
    assert_close_dist(refl.cpu(), want_r, "reflect dir", median=1e-6, frac=0.0, p99=1e-5)
    comp, index = compact_rows(sec, m)
    assert torch.equal(comp.cpu(), sec.cpu()[wm.bool()])  # stable order == boolean-mask indexing
    child = torch.rand(comp.shape[0], 3, generator=gen)
    cdepth = torch.rand(comp.shape[0], generator=gen)
    base = torch.rand(n, 3, generator=gen)
    rgb, rr, dr = blend_reflection(base.cuda(), m, child.cuda(), cdepth.cuda(), index)
    mb = wm.bool()
    refl_full = base.clone()
    refl_full[mb] = child
    m3 = mb.float()[:, None]
    assert torch.allclose(rgb.cpu(), m3 * refl_full + (1 - m3) * base, atol=1e-7)
    want_rr = torch.zeros(n, 3)
    want_rr[mb] = child
    assert torch.equal(rr.cpu(), want_rr)
    want_dr = torch.zeros(n)
    want_dr[mb] = cdepth
    assert torch.equal(dr.cpu(), want_dr)
    # no mirror at all
    z = torch.zeros(100, device="cuda")
    _, _, flag0 = reflect_rays(rays[:100].cuda(), xs[:100].cuda(), nrm[:100].cuda(), z)
    assert int(flag0.item()) == 0
    comp0, _ = compact_rows(sec[:100].contiguous(), z)
    assert comp0.shape[0] == 0


@pytest.mark.parametrize("levels", [1, 2])
def test_recursive_render_vs_oracle(oracle, mm, params, levels):
    """Eval-semantics Whitted recursion (1 and 2 bounces) against the oracle's restatement of eval.py."""
    from mirror_nerf_b200.synthetic import random_rays
    from mirror_nerf_b200.trace import render_rays_recursive
    models, emb = mm
    rays = random_rays(96, seed=12)
    with torch.no_grad():
        got = render_rays_recursive(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False,
                                    max_recursive_level=levels)
        fn = lambda r: oracle.render_rays(params, r, 64, False, 0, 0, 128, 32768, False, test_time=True,
                                          compute_normal=False)
        want = oracle.trace_eval(fn, rays, levels)
        # The synthetic field is chaotic (sigma head x40, 2^9 frequencies): the REFERENCE's own bounce output moves
        # when its input rays move by one ulp.  That self-sensitivity is the noise floor our deviation is judged by.
        rays_ulp = rays.clone()
        rays_ulp[:, :6] = torch.nextafter(rays_ulp[:, :6], torch.full_like(rays_ulp[:, :6], 10.0))
        want_ulp = oracle.trace_eval(fn, rays_ulp, levels)
    assert set(want) <= set(got), sorted(set(want) - set(got))
    flips = float((got["mirror_mask_fine"].cpu() != want["mirror_mask_fine"]).float().mean())
    assert flips <= 0.03
    # level-0 quantities are not amplified by the bounce: tight
    assert_close_dist(got["rgb_fine_direct"].cpu(), want["rgb_fine_direct"], f"bounce{levels} direct")
    assert_close_dist(got["depth_fine"].cpu(), want["depth_fine"], f"bounce{levels} depth")
    same = (got["mirror_mask_fine"].cpu() == want["mirror_mask_fine"]) & \
           (want_ulp["mirror_mask_fine"] == want["mirror_mask_fine"])
    for k in ("rgb_fine", "rgb_fine_reflect", "depth_fine_reflect"):
        ours = err_stats(got[k].cpu()[same], want[k][same])
        floor = err_stats(want_ulp[k][same], want[k][same])
        msg = fmt_stats(f"bounce{levels} {k} ours", ours) + " || " + fmt_stats("reference +1ulp", floor)
        assert ours["median"] <= max(2e-4, 2.0 * floor["median"]), msg
        assert ours["frac"] <= max(0.03, 1.5 * floor["frac"] + 0.03), msg


@pytest.mark.parametrize("levels,only_mirror", [(1, None), (2, None), (2, True)])
def test_recursion_logic_on_smooth_field(oracle, levels, only_mirror):
    """The recursion driver itself (threshold, reflect, stable compaction, recursion depth, blend) against the oracle's
    restatement of eval.py / train.py, with a smooth analytic stand-in for render_rays so nothing is chaotic."""
    from mirror_nerf_b200.trace import render_rays_recursive

    def fake(r):
        o, d = r[:, :3], r[:, 3:6]
        depth = 1.0 + 0.5 * torch.sin(o.sum(-1)) ** 2
        return {"rgb_fine": 0.5 + 0.5 * torch.sin(o * 1.3 + d * 0.7),
                "depth_fine": depth,
                "mirror_mask_fine": 0.5 + 0.5 * torch.sin(3.0 * o[:, 0] + d[:, 1]),
                "surface_normal_fine": torch.stack([torch.cos(o[:, 1]), torch.sin(o[:, 2]), 0.3 + d[:, 0] ** 2], -1),
                "x_surface_fine": o + d * depth[:, None]}

    gen = torch.Generator().manual_seed(31)
    rays = torch.cat([torch.randn(3000, 6, generator=gen), torch.full((3000, 1), 0.05), torch.full((3000, 1), 8.0)], 1)
    got = render_rays_recursive(None, None, rays.cuda(), 64, False, 0, 0, 128, 32768, False,
                                max_recursive_level=levels, only_trace_rays_in_mirrors=only_mirror,
                                render_fn=lambda r: {k: v.contiguous() for k, v in fake(r).items()})
    if only_mirror is None:
        want = oracle.trace_eval(fake, rays, levels)
    else:  # train.py semantics: compact at every level
        def trace_train(r, level=0):
            res = fake(r)
            m = res["mirror_mask_fine"]
            m[m > 0.5] = 1
            m[m < 0.5] = 0
            mb = m.bool()
            if bool(mb.any()) and level < levels:
                sec, _ = oracle.reflect_rays(r, res["x_surface_fine"], res["surface_normal_fine"])
                sub = trace_train(sec[mb], level + 1)
                refl = res["rgb_fine"].clone()
                refl[mb] = sub["rgb_fine"]
                m3 = mb.float()[:, None]
                res["rgb_fine"] = m3 * refl + (1 - m3) * res["rgb_fine"]
            return res
        want = trace_train(rays)
    assert torch.equal(got["mirror_mask_fine"].cpu(), want["mirror_mask_fine"])
    assert_close_dist(got["rgb_fine"].cpu(), want["rgb_fine"], "logic rgb", median=1e-6, frac=0.0, p99=1e-4)
    if only_mirror is None:
        assert_close_dist(got["rgb_fine_reflect"].cpu(), want["rgb_fine_reflect"], "logic reflect", median=1e-6,
                          frac=0.0, p99=1e-4)
        assert_close_dist(got["depth_fine_reflect"].cpu(), want["depth_fine_reflect"], "logic depth_reflect",
                          median=1e-6, frac=0.0, p99=1e-4)


# ---------------------------------------------------------------------------------------------- section 8f rows 3-4
def test_roughness_cone_logic_on_smooth_field(oracle):
    """--app_control_mirror_roughness: jittered normals, T extra reflections of the mirror rays averaged (eval.py:623-674)."""
    from mirror_nerf_b200.trace import render_rays_recursive

    def fake(r):
        o, d = r[:, :3], r[:, 3:6]
        depth = 1.0 + 0.5 * torch.sin(o.sum(-1)) ** 2
        return {"rgb_fine": 0.5 + 0.5 * torch.sin(o * 1.3 + d * 0.7), "depth_fine": depth,
                "mirror_mask_fine": 0.5 + 0.5 * torch.sin(3.0 * o[:, 0] + d[:, 1]),
                "surface_normal_fine": torch.stack([torch.cos(o[:, 1]), torch.sin(o[:, 2]), 0.3 + d[:, 0] ** 2], -1),
                "x_surface_fine": o + d * depth[:, None]}

    gen = torch.Generator().manual_seed(41)
    n, T_extra = 2000, 3
    rays = torch.cat([torch.randn(n, 6, generator=gen), torch.full((n, 1), 0.05), torch.full((n, 1), 8.0)], 1)
    noises = [0.05 * torch.randn(n, 3, generator=gen) for _ in range(T_extra + 1)]
    got = render_rays_recursive(None, None, rays.cuda(), 64, False, 0, 0, 128, 32768, False, max_recursive_level=1,
                                render_fn=lambda r: {k: v.contiguous() for k, v in fake(r).items()},
                                normal_noise_std=0.05, trace_ray_times=T_extra, normal_noises=[z.cuda() for z in noises])
    want = oracle.trace_eval(fake, rays, 1, normal_noises=noises, trace_ray_times=T_extra)
    assert torch.equal(got["mirror_mask_fine"].cpu(), want["mirror_mask_fine"])
    assert_close_dist(got["rgb_fine"].cpu(), want["rgb_fine"], "roughness rgb", median=1e-6, frac=0.0, p99=1e-4)


def test_generate_rays_matches_reference_camera():
    """Device ray generation vs the CPU restatement of ray_utils.get_ray_directions/get_rays (blender.py:158-168)."""
    import math
    from mirror_nerf_b200.ray_utils import focal_from_fov, generate_rays
    from mirror_nerf_b200.synthetic import camera_rays
    a = 0.4
    c2w = torch.tensor([[math.cos(a), 0.0, math.sin(a), 1.0], [0.0, 1.0, 0.0, -0.5], [-math.sin(a), 0.0, math.cos(a), 2.5]])
    for H, W in ((400, 400), (300, 400), (7, 5)):
        fov = 0.6911112070083618
        want = camera_rays(H, W, fov, c2w, near=0.05, far=8.0)
        got = generate_rays(H, W, focal_from_fov(W, fov), c2w, 0.05, 8.0).cpu()
        assert got.shape == want.shape
        assert torch.equal(got[:, [0, 1, 2, 6, 7]], want[:, [0, 1, 2, 6, 7]])
        assert float((got[:, 3:6] - want[:, 3:6]).abs().max()) <= 2e-7  # matmul summation order / FMA contraction


def test_room_train_mode_forward(oracle):
    """Train-mode forward (test_time=False, compute_normal=True, perturb = noise_std = 1 with the draws replayed) on the SCENE-LIKE
    field: both passes full, analytic normals from the tcgen05 reverse chain, vs the oracle on the same rays and draws."""
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.room_scene import room_pose
    from mirror_nerf_b200.synthetic import camera_rays
    from util import room_state_dicts
    sds = room_state_dicts()
    models = {}
    for k, sd in sds.items():
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        models[k] = m.cuda().eval()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    allrays = camera_rays(200, 200, c2w=room_pose(3), near=0.05, far=12.0)
    n = 256
    rays = allrays[torch.linspace(0, allrays.shape[0] - 1, n).long()].contiguous()
    g = torch.Generator().manual_seed(17)
    rng = {"perturb_u": torch.rand(n, 64, generator=g), "noise_coarse": torch.randn(n, 64, generator=g),
           "u_pdf": torch.rand(n, 128, generator=g), "noise_fine": torch.randn(n, 192, generator=g)}
    args = (64, False, 1.0, 1.0, 128, 32768, False)
    want = oracle.render_rays(sds, rays, *args, test_time=False, compute_normal=True, rng=rng)
    with torch.no_grad():
        got = render_rays(models, emb, rays.cuda(), *args, test_time=False, compute_normal=True, rng=rng)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    assert torch.equal(got["z_vals_coarse"].cpu(), want["z_vals_coarse"])
    for k in sorted(got):
        per_sample = got[k].dim() >= 2 and got[k].shape[1] in (64, 192)
        # per-ray outputs: tight; per-sample tensors (weights, normals of points in empty space) carry the field's own
        # conditioning: a looser tail
        assert_close_dist(got[k].cpu(), want[k].detach(), f"room train-mode {k}", median=5e-5, frac=0.05 if per_sample else 0.02)


def test_sharded_render_equals_unsharded_bitwise(mm):
    """rays are independent units, so a ray's result must not depend on which rank's shard it lands in --
    rendering the tile-aligned shards of mirror_nerf_b200.parallel.shard_bounds one after the other (what N ranks do) gives
    bit-identical outputs to rendering the whole batch."""
    from mirror_nerf_b200.parallel import shard_bounds
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import random_rays
    models, emb = mm
    n = 1000
    rays = random_rays(n, seed=21).cuda()
    args = (64, False, 0, 0, 128, 32768, False)
    with torch.no_grad():
        whole = render_rays(models, emb, rays, *args, test_time=True, compute_normal=False)
        for world in (2, 3, 8):
            parts = []
            for rank in range(world):
                lo, hi = shard_bounds(n, rank, world)
                if hi > lo:
                    parts.append(render_rays(models, emb, rays[lo:hi].contiguous(), *args, test_time=True, compute_normal=False))
            for k in ("rgb_fine", "depth_fine", "mirror_mask_fine", "weights_fine", "z_vals_fine", "surface_normal_fine"):
                got = torch.cat([p[k] for p in parts], 0)
                assert torch.equal(got, whole[k]), (world, k, float((got - whole[k]).abs().max()))


def test_room_scene_parity_and_psnr(oracle):
    """The north-star bar on a SCENE-LIKE field (tests/golden/room_field.npz: our training path fitted to the analytic room of
    mirror_nerf_b200/room_scene.py): one bounce with eval semantics, ours (tc3) vs the oracle on the same rays --
    rgb / depth within 1e-3 (relative to max(|want|, rms)) on >= 99 % of the rays, and PSNR against the analytic ground truth
    within 0.05 dB of the reference's."""
    import math
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    from mirror_nerf_b200.room_scene import room_pose, trace_room
    from mirror_nerf_b200.synthetic import camera_rays
    from mirror_nerf_b200.trace import render_rays_recursive
    from util import room_state_dicts
    sds = room_state_dicts()
    models = {}
    for k, sd in sds.items():
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        models[k] = m.cuda().eval()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    allrays = camera_rays(200, 200, c2w=room_pose(1), near=0.05, far=12.0)
    rays = allrays[torch.linspace(0, allrays.shape[0] - 1, 1536).long()].contiguous()
    gt, gt_mask, _ = trace_room(rays)
    fn = lambda r: oracle.render_rays(sds, r, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
    with torch.no_grad():
        want = oracle.trace_eval(fn, rays, 1)
        got = render_rays_recursive(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False, max_recursive_level=1)
    for k in ("rgb_fine", "depth_fine", "opacity_fine"):
        s = err_stats(got[k].cpu(), want[k], name=f"room tc3 {k}")
        assert s["median"] <= 1e-5 and s["frac"] <= 0.01, fmt_stats(k, s)
    assert float((got["mirror_mask_fine"].cpu() != want["mirror_mask_fine"]).float().mean()) <= 0.002  # thresholded masks
    psnr = lambda x: -10 * math.log10(float(((x - gt) ** 2).mean()))
    p_ours, p_ref = psnr(got["rgb_fine"].cpu()), psnr(want["rgb_fine"])
    print(f"room scene: PSNR ours {p_ours:.4f} dB, reference {p_ref:.4f} dB, mirror fraction {float(gt_mask.mean()):.3f}")
    assert p_ref > 18.0, "fixture should be a fitted scene"
    assert abs(p_ours - p_ref) <= 0.05
    # the single-pass speed mode on the same scene (informational + the PSNR bar only: its per-ray errors are ~1e-3)
    with torch.no_grad():
        fast = render_rays_recursive(models, emb, rays.cuda(), 64, False, 0, 0, 128, 32768, False, max_recursive_level=1,
                                     field_impl="tc1")
    s1 = err_stats(fast["rgb_fine"].cpu(), want["rgb_fine"])
    print(f"room scene, tc1: PSNR {psnr(fast['rgb_fine'].cpu()):.4f} dB; " + fmt_stats("rgb_fine", s1))
    assert abs(psnr(fast["rgb_fine"].cpu()) - p_ref) <= 0.05
