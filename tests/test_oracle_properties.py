"""Property tests (hypothesis) of the oracle's samplers on the edge cases SURVEY.md appendix A lists: all-zero weights, single
spikes, u = 0 / u = 1, ties.  CPU only; these pin the restatement the GPU kernels are compared against."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import mirror_nerf_oracle as O


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), kind=st.sampled_from(["zero", "spike", "two_spikes", "tiny", "random"]),
       n_imp=st.sampled_from([1, 16, 128]))
def test_sample_pdf_is_monotone_and_inside_the_bins(seed, kind, n_imp):
    g = np.random.Generator(np.random.PCG64(seed))
    n, nb = 4, 63
    z = np.sort(g.uniform(0.05, 8.0, size=(n, nb)).astype(np.float32), axis=1)
    w = np.zeros((n, nb - 1), np.float32)
    if kind == "spike":
        w[np.arange(n), g.integers(0, nb - 1, n)] = 1.0
    elif kind == "two_spikes":
        w[:, 0] = 0.7
        w[:, -1] = 0.3
    elif kind == "tiny":
        w[:] = 1e-7
    elif kind == "random":
        w = g.uniform(0, 1, size=(n, nb - 1)).astype(np.float32) ** 8
    bins, wt = torch.from_numpy(z), torch.from_numpy(w)
    s, inds, cdf = O.sample_pdf(bins, wt, n_imp, det=True, return_inds=True)
    assert torch.isfinite(s).all()
    assert bool((s[:, 1:] >= s[:, :-1]).all())                      # u is increasing -> samples are non-decreasing
    assert bool((s >= bins[:, :1]).all()) and bool((s <= bins[:, -1:]).all())
    assert int(inds.min()) >= 1 and int(inds.max()) <= nb           # searchsorted(right=True) of u in [0,1] against cdf[0] = 0
    assert bool((cdf[:, 1:] >= cdf[:, :-1]).all()) and float(cdf[:, 0].abs().max()) == 0.0
    assert torch.allclose(cdf[:, -1], torch.ones(n), atol=1e-5)
    assert torch.equal(s[:, 0], bins[:, 0])                          # u = 0 -> first bin edge (appendix A)


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), perturb=st.sampled_from([0.0, 0.5, 1.0]), use_disp=st.booleans())
def test_coarse_depths_are_sorted_and_inside_near_far(seed, perturb, use_disp):
    g = torch.Generator().manual_seed(seed)
    n = 5
    rays = torch.zeros(n, 8)
    rays[:, 6] = 0.05 + torch.rand(n, generator=g)
    rays[:, 7] = rays[:, 6] + 0.5 + 5 * torch.rand(n, generator=g)
    u = torch.rand(n, 64, generator=g)
    z = O.coarse_z_vals(rays, 64, use_disp, perturb, u if perturb > 0 else None)
    assert z.shape == (n, 64)
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert bool((z >= rays[:, 6:7] - 1e-6).all()) and bool((z <= rays[:, 7:8] + 1e-5).all())


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), white_back=st.booleans(), S=st.sampled_from([2, 3, 33, 64]))
def test_composite_weights_form_a_sub_partition_of_unity(seed, white_back, S):
    # S >= 2: with a single sample the reference's own `ones_like(deltas[:, :1])` is empty and every output degenerates
    g = torch.Generator().manual_seed(seed)
    n = 6
    z = torch.sort(torch.rand(n, S, generator=g) * 4 + 0.1, -1)[0]
    sig = torch.randn(n, S, generator=g) * 30
    rgb = torch.rand(n, S, 3, generator=g)
    res = {}
    O.composite(res, "fine", z, sig, rgb, None, None, None, noise_std=0.0, white_back=white_back, weights_only=False,
                noise=torch.zeros(n, S))
    w = res["weights_fine"]
    assert bool((w >= 0).all()) and bool((res["opacity_fine"] <= 1 + 1e-5).all())
    # last interval is 1e10 wide: opacity is 1 exactly when the last sample is dense, otherwise < 1
    dense_last = sig[:, -1] > 0
    assert torch.allclose(res["opacity_fine"][dense_last], torch.ones(int(dense_last.sum())), atol=1e-4)
    assert bool((res["rgb_fine"] >= -1e-6).all()) and bool((res["rgb_fine"] <= 1 + 1e-4).all())


def test_z_fine_hook_reproduces_the_sampled_run():
    """rng["z_fine"] (the gradient tests' hook: evaluate the oracle at the fine depths of the implementation under test) changes
    nothing when it is fed the oracle's own depths, and the depths carry no gradient (R/models/rendering.py:346-349 detaches)."""
    import torch
    from oracle import mirror_nerf_oracle as O
    from mirror_nerf_b200.synthetic import random_rays, scene_state_dicts
    sds = scene_state_dicts()
    rays = random_rays(5, seed=3)
    g = torch.Generator().manual_seed(1)
    rng = {"perturb_u": torch.rand(5, 16, generator=g), "noise_coarse": torch.randn(5, 16, generator=g),
           "u_pdf": torch.rand(5, 24, generator=g), "noise_fine": torch.randn(5, 40, generator=g)}
    args = (16, False, 1.0, 1.0, 24, 32768, False)
    with torch.no_grad():
        a = O.render_rays(sds, rays, *args, rng=rng, test_time=False)
        b = O.render_rays(sds, rays, *args, rng=dict(rng, z_fine=a["z_vals_fine"]), test_time=False)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert not a["z_vals_fine"].requires_grad
