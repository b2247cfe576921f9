"""The oracle's restatements of the CALLERS of render_rays and of the helpers either side of the path (SURVEY.md section 8f rows
1, 3, 4) against the unmodified reference (fixture `golden`, see conftest.py / tests/golden/make_golden_recursion.py):

* oracle.trace_eval   vs  R/eval.py::batched_inference (:114-740), 1 and 2 bounces, and --app_control_mirror_roughness
* oracle.trace_train  vs  R/train.py::NeRFSystem.render_rays_chunk_recursively (:129-348), outputs AND parameter gradients
* mirror_nerf_b200.checkpoint           vs  R/utils/__init__.py:109-136
* mirror_nerf_b200.synthetic.camera_rays vs R/datasets/ray_utils.py:6-53

"live" = the reference runs in this process -> bit-identical; "file" = committed vectors from another host CPU -> host-tolerant
bounds written below (the trained room field is well conditioned: the bounds are much tighter than for the adversarial field)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mirror_nerf_oracle as O
from util import err_stats, room_state_dicts

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def T(x):
    return torch.from_numpy(np.asarray(x))


def case(g, name):
    return {k[len(name) + 1:]: v for k, v in g.items() if k.startswith(name + "/")}


def same(a, want, name, exact, med=1e-5, frac=0.02):
    a = a.detach()
    b = T(want)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype == torch.bool or exact:
        assert torch.equal(a, b.to(a.dtype)), (name, float((a.double() - b.double()).abs().max()))
    else:
        s = err_stats(a, b)
        assert s["median"] <= med and s["frac"] <= frac, (name, s)


@pytest.mark.parametrize("name", ["l1", "l2", "rough_l2"])
def test_trace_eval_vs_reference_batched_inference(golden, name):
    import make_golden_recursion as M
    n, mirror_only, levels, rough, T_extra, std, seed = M.EVAL_CASES[name]
    c = case(golden("recursion_eval"), name)
    rays = T(c["rays"])
    sds = room_state_dicts()
    fn = lambda r: O.render_rays(sds, r, 64, False, 0, 0, 128, 32768, False, test_time=True, compute_normal=False)
    torch.manual_seed(1000 + seed)   # the reference draws randn_like(normal) * std from the global stream (R/eval.py:506-511)
    with torch.no_grad():
        got = O.trace_eval(fn, rays, levels, trace_ray_times=T_extra,
                           noise_fn=(lambda k: torch.randn(k, 3) * std) if rough else None)
    want = {k[4:]: v for k, v in c.items() if k.startswith("out/")}
    assert set(want) <= set(got), sorted(set(want) - set(got))
    assert set(got) - set(want) <= {"rgb_fine_direct"}
    assert torch.equal(got["mirror_mask_fine"], T(want["mirror_mask_fine"])), "thresholded masks"
    for k in want:
        same(got[k], want[k], f"{name}/{k}", golden.exact)


@pytest.mark.parametrize("name", ["gt_only_mirror", "pred_all_rays_l2"])
def test_trace_train_vs_reference_nerf_system(golden, name):
    import make_golden_recursion as M
    cfg = M.TRAIN_CASES[name]
    c = case(golden("recursion_train"), name)
    rays, gt = T(c["rays"]), T(c["gt_mask"])
    sds = room_state_dicts()
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in sds.items()}
    fn = lambda r: O.render_rays(params, r, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False, compute_normal=True,
                                 mirror_mask=gt)
    torch.manual_seed(M.TRAIN_SEED)
    got = O.trace_train(fn, rays, gt, cfg["levels"], only_trace_rays_in_mirrors=cfg["only"], is_eval=cfg["is_eval"],
                        detach_normal_in_reflection=cfg["detach_normal"], detach_ref_color=cfg["detach_ref"])
    want = {k[4:]: v for k, v in c.items() if k.startswith("out/")}
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    for k in want:
        # perturbed samples + sigma noise + autograd normals of 16-24 rays: the reference itself moves host to host
        same(got[k], want[k], f"{name}/{k}", golden.exact, med=1e-4, frac=0.10)
    loss = M.train_loss(got)
    loss.backward()
    if golden.exact:
        assert torch.equal(loss.detach(), T(c["loss"]))
    else:
        assert abs(float(loss.detach()) - float(c["loss"])) <= 1e-4 * abs(float(c["loss"]))
    for tag in ("coarse", "fine"):
        for k, t in params[tag].items():
            gr = t.grad.flatten()[::13] if t.grad.numel() > 4096 else t.grad
            wg, wn = T(c[f"grad/{tag}/{k}"]), float(c[f"gradnorm/{tag}/{k}"])
            if golden.exact:
                assert torch.equal(gr.detach(), wg), (tag, k)
                continue
            a, b = gr.detach().double().flatten(), wg.double().flatten()
            if float(b.norm()) > 0:
                cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
                assert cos > 0.999, (tag, k, cos)
            assert abs(float(t.grad.norm()) - wn) <= 2e-2 * wn + 1e-9, (tag, k)


def test_checkpoint_import_vs_reference(golden, tmp_path):
    """extract_model_state_dict / load_ckpt (R/utils/__init__.py:109-136) on a pytorch-lightning style file."""
    import make_golden_recursion as M
    from mirror_nerf_b200.checkpoint import extract_model_state_dict, load_ckpt
    from mirror_nerf_b200.mirror_nerf import MirrorNeRF
    c = case(golden("helpers"), "ckpt")
    sds = room_state_dicts()
    path = M.write_pl_checkpoint(sds, str(tmp_path / "pl.ckpt"))
    ext = extract_model_state_dict(path, "nerf_fine", prefixes_to_ignore=["normal_net"])
    assert sorted(ext.keys()) == [str(k) for k in c["extract_keys"]]
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    m.load_state_dict(sds["coarse"])
    load_ckpt(m, path, "nerf_fine", prefixes_to_ignore=["normal_net"])
    sd = m.state_dict()
    loaded = {k[len("loaded/"):]: v for k, v in c.items() if k.startswith("loaded/")}
    assert set(loaded) == set(sd)
    for k, v in sd.items():
        assert torch.equal(v[..., :5] if v.dim() == 2 else v[:5], T(loaded[k])), k
    load_ckpt(m, "", "nerf_fine")   # empty path: silent no-op, as the reference
    with pytest.raises(AssertionError):
        load_ckpt(m, path, "no_such_model")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_camera_rays_vs_reference_ray_utils(golden, tag):
    """get_ray_directions + get_rays (R/datasets/ray_utils.py:6-53): the CPU restatement that the device kernel is tested against."""
    import math
    from mirror_nerf_b200.synthetic import camera_rays
    c = case(golden("helpers"), f"rays_{tag}")
    H, W, focal = int(c["HWf"][0]), int(c["HWf"][1]), float(c["HWf"][2])
    fov = 2.0 * math.atan(0.5 * W / focal)
    got = camera_rays(H, W, fov, T(c["c2w"]), near=0.05, far=8.0)
    assert torch.equal(got[:, 0:3], T(c["rays_o"]))
    if golden.exact:
        assert float((got[:, 3:6] - T(c["rays_d"])).abs().max()) <= 1.2e-7   # focal goes through atan/tan once more: 1 ulp
    else:
        assert float((got[:, 3:6] - T(c["rays_d"])).abs().max()) <= 5e-7
