"""The CPU oracle (oracle/mirror_nerf_oracle.py) against the unmodified reference, twice (fixture `golden`):

* "live": the reference is imported from /root/reference and run in this process -> the oracle must be BIT-IDENTICAL
  (same ATen kernels, same host).  Skipped where the reference tree does not exist (the GPU box).
* "file": the committed vectors of tests/golden/*.npz.  They were produced on another host CPU; the reference itself
  moves by ~5e-7 per point between hosts (different BLAS kernels) and the sharp synthetic field amplifies that along a
  ray (observed reference-vs-reference: up to 1.7e-3 on single `weights_fine` entries).  Per-point field outputs:
  1e-5 rel / 2e-6 abs.  Rendered outputs: median relative error <= 1e-5 and at most 5 % of a tensor's entries off
  by more than 1e-3 (relative to max(|want|, rms)).  Indices / CDF of sample_pdf on given weights are exact."""
import numpy as np
import pytest
import torch

from mirror_nerf_b200.synthetic import scene_state_dicts
from oracle import mirror_nerf_oracle as O
from util import err_stats


def T(x):
    return torch.from_numpy(np.asarray(x))


def close(a, b, name, rtol=1e-5, atol=2e-6, exact=False, rendered=False, med=1e-5, frac=0.05):
    a = a.detach() if isinstance(a, torch.Tensor) else T(a)
    b = T(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if exact:
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    elif rendered:
        s = err_stats(a, b)
        assert s["median"] <= med and s["frac"] <= frac, (name, s)
    else:
        assert torch.allclose(a, b, rtol=rtol, atol=atol), (name, float((a - b).abs().max()))


@pytest.fixture(scope="module")
def params():
    return scene_state_dicts()


def test_embedding(golden):
    g = golden("field")
    assert torch.equal(O.embed(T(g["xyz"]), 10), T(g["pe_xyz"]))
    assert torch.equal(O.embed(T(g["dir"]), 4), T(g["pe_dir"]))


def test_field_forward(golden, params):
    g = golden("field")
    x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1)
    with torch.no_grad():
        o = O.field_forward(params["fine"], x.clone(), compute_normal=False, sigma_only=False)
    for k in ("sigma", "geo_feat", "pred_normal", "rgb", "is_mirror"):
        close(o[k], g["full_" + k], k, exact=golden.exact)
    with torch.no_grad():
        o = O.field_forward(params["fine"], T(g["xyz"]).clone(), compute_normal=False, sigma_only=True)
    assert set(o) == {"sigma", "geo_feat", "pred_normal"}
    close(o["sigma"], g["sigonly_sigma"], "sigonly_sigma", exact=golden.exact)
    close(o["pred_normal"], g["sigonly_pred_normal"], "sigonly_pred_normal", exact=golden.exact)
    o = O.field_forward(params["fine"], x.clone(), compute_normal=True, sigma_only=False)
    for k in ("sigma", "normal", "pred_normal", "rgb", "is_mirror"):
        # the autograd normal is a normalised gradient of a sharp field: BLAS kernel choice on another host CPU moves it
        # by a few 1e-6 (observed 2.2e-6), so it gets a looser absolute tolerance than the forward outputs
        close(o[k], g["grad_" + k], "grad_" + k, atol=2e-5 if k == "normal" else 2e-6, exact=golden.exact)


def test_explicit_normal_chain_matches_autograd(golden, params):
    g = golden("field")
    xyz = T(g["xyz"])
    with torch.no_grad():
        gx = O.analytic_normal_explicit(params["fine"], xyz)
    n = O.l2_normalize(-gx)
    ref = T(g["grad_normal"])
    # autograd and the explicit chain sum in different orders: compare directions
    cos = (n * ref).sum(-1)
    assert float(cos.min()) > 1 - 1e-4


def test_sample_pdf(golden):
    g = golden("sample_pdf")
    bins, w = T(g["bins"]), T(g["weights"])
    s, inds, cdf = O.sample_pdf(bins, w, 128, det=True, return_inds=True)
    assert torch.equal(inds, T(g["inds_det"]))
    close(cdf, g["cdf"], "cdf", atol=0, rtol=0)
    close(s, g["det"], "det", exact=golden.exact)
    s, inds, _ = O.sample_pdf(bins, w, 128, det=False, u=T(g["u"]), return_inds=True)
    assert torch.equal(inds, T(g["inds_rnd"]))
    close(s, g["rnd"], "rnd", exact=golden.exact)
    assert bool((s[:, :] >= bins[:, :1]).all()) and bool((s <= bins[:, -1:]).all())


def test_render_eval(golden, params):
    g = golden("render_eval")
    with torch.no_grad():
        r = O.render_rays(params, T(g["rays"]), 64, False, 0, 0, 128, 32768, False, test_time=True,
                          compute_normal=False)
    assert set(r) == set(g) - {"rays"}
    for k in r:
        close(r[k], g[k], k, exact=golden.exact, rendered=True)


VARIANTS = {
    "white_back": dict(args=(64, False, 0, 0, 128, 32768, True), kw=dict(test_time=True), fine=True),
    "use_disp": dict(args=(64, True, 0, 0, 128, 32768, False), kw=dict(test_time=True), fine=True),
    "coarse_only": dict(args=(64, False, 0, 0, 0, 32768, False), kw=dict(test_time=True), fine=False),
    "one_field": dict(args=(64, False, 0, 0, 128, 32768, False),
                      kw=dict(test_time=True, only_one_field=True, current_epoch=3), fine=False),
    "train_nonormal": dict(args=(64, False, 0, 0, 128, 32768, False), kw=dict(test_time=False), fine=True),
    "s32_i16": dict(args=(32, False, 0, 0, 16, 1000, False), kw=dict(test_time=True), fine=True),
}


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_render_variants(golden, params, tag):
    g = golden("render_variants")
    v = VARIANTS[tag]
    p = params if v["fine"] else {"coarse": params["coarse"]}
    with torch.no_grad():
        r = O.render_rays(p, T(g["rays"]), *v["args"], compute_normal=False, **v["kw"])
    want = {k.split("/", 1)[1]: a for k, a in g.items() if k.startswith(tag + "/")}
    assert set(r) == set(want)
    for k in r:
        close(r[k], want[k], f"{tag}/{k}", exact=golden.exact, rendered=True)


def test_render_train_with_grads(golden):
    g = golden("render_train")
    params = scene_state_dicts()
    for p in params.values():
        for t in p.values():
            t.requires_grad_(True)
    rng = {k.split("/", 1)[1]: T(a) for k, a in g.items() if k.startswith("rng/")}
    r = O.render_rays(params, T(g["rays"]), 64, False, 1.0, 1.0, 128, 32768, False, test_time=False,
                      compute_normal=True, rng=rng)
    want = {k.split("/", 1)[1]: a for k, a in g.items() if k.startswith("out/")}
    assert set(r) == set(want)
    for k in r:
        # 8 rays, perturbed samples, normals = normalised autograd gradients of the sharp field: the host-to-host
        # movement of the reference itself reaches 2e-3 on a few of the 24 composited-normal entries
        close(r[k], want[k], k, exact=golden.exact, rendered=True, med=1e-4, frac=0.25)
    loss = sum((r[f"rgb_{t}"] ** 2).sum() + r[f"mirror_mask_{t}"].sum() + 0.1 * r[f"normal_dif_{t}"].sum()
               + 0.01 * (r[f"depth_{t}"]).sum() for t in ("coarse", "fine"))
    loss.backward()
    close(loss, g["loss"], "loss", rtol=1e-4, exact=golden.exact)
    for tag in ("coarse", "fine"):
        for k, t in params[tag].items():
            gr = t.grad.flatten()[::13] if t.grad.numel() > 4096 else t.grad
            want_g = T(g[f"grad/{tag}/{k}"])
            if golden.exact:
                assert torch.equal(gr.detach(), want_g), k
                assert torch.equal(t.grad.norm(), T(g[f"gradnorm/{tag}/{k}"])), k
                continue
            # host-tolerant: direction of the (subsampled) gradient and its norm
            a, b = gr.detach().double().flatten(), want_g.double().flatten()
            cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
            assert cos > 0.999, (k, cos)
            gn, wn = float(t.grad.norm()), float(T(g[f"gradnorm/{tag}/{k}"]))
            assert abs(gn - wn) <= 2e-2 * wn + 1e-9, (k, gn, wn)


def test_hashgrid_oracle_regression_vectors():
    """tests/golden/hash_field.npz (tests/golden/make_golden_hash.py): REGRESSION vectors of the hash-grid restatement, written by
    the oracle itself -- the reference's encoder (tinycudann) cannot run here, the hash-grid path stays parity-unpinned.  Level
    table exact; encoded features / field outputs 1e-5 rel + 2e-6 abs (host BLAS); analytic normals by cosine; training
    gradients (incl. the double backward through the analytic normal) to 1e-4 of the tensor scale."""
    import os
    from oracle import hashgrid_oracle as H
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_field.npz"))
    for bound in (1.0, 2.0):
        tag = f"b{int(bound)}"
        lv, total = H.level_table(bound)
        assert np.array_equal(np.array([[s, r, o, n] for s, r, o, n in lv], np.float64), g[f"levels_{tag}"])
        assert total == int(g[f"total_{tag}"][0])
        sd = H.make_state_dict(3, bound=bound)
        x = T(g[f"x_{tag}"])
        enc = H.hashgrid_encode(sd["encoder.params"], (x[:, :3] + bound) / (2 * bound), bound)
        close(enc, g[f"enc_{tag}"], "enc", rtol=1e-5, atol=2e-6)
        o = H.field_forward(sd, x, bound=bound, compute_normal=True)
        for k in ("sigma", "rgb", "pred_normal", "is_mirror"):
            close(o[k], g[f"{k}_{tag}"], k, rtol=1e-4, atol=1e-5)
        cos = (o["normal"].detach() * T(g[f"normal_{tag}"])).sum(-1)
        assert float(cos.min()) > 1 - 1e-4, float(cos.min())
    sd = H.make_state_dict(3, sigma_scale=4.0)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x, c = T(g["train_x"]), T(g["train_c"])
    o = H.field_forward(p, x, compute_normal=True)
    loss = (o["sigma"][:, 0] * c[:, 0]).sum() + (o["rgb"] * c[:, 1:4]).sum() + (o["is_mirror"][:, 0] * c[:, 4]).sum() + \
        (o["pred_normal"] * c[:, 5:8]).sum() + (o["normal"] * c[:, 8:11]).sum()
    loss.backward()
    for k, v in p.items():
        if k == "encoder.params":
            got = v.grad[T(g["grad_table_idx"])]
            want = T(g["grad_table_val"])
            assert int(torch.count_nonzero(v.grad)) == want.numel()
        else:
            got, want = v.grad, T(g["grad_" + k])
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()), k
