"""Golden vectors for the CALLERS of render_rays and the data helpers either side of the hot path (SURVEY.md section 8f rows
1, 3, 4), produced by running the UNMODIFIED reference (needs /root/reference; build container only):

    python tests/golden/make_golden_recursion.py

* ``R/train.py::NeRFSystem.render_rays_chunk_recursively`` (:129-348)   -> recursion_train.npz
* ``R/eval.py::batched_inference`` (:114-740), plain and with ``--app_control_mirror_roughness`` -> recursion_eval.npz
* ``R/utils/__init__.py::extract_model_state_dict / load_ckpt`` (:109-136)                        -> helpers.npz
* ``R/datasets/ray_utils.py::get_ray_directions / get_rays`` (:6-53)                              -> helpers.npz

``train.py`` / ``eval.py`` import pytorch_lightning, kornia, imageio and torch_optimizer, none of which exist offline and none
of which the recursion touches: they are replaced by empty stand-ins in ``sys.modules`` (``LightningModule`` = plain object;
``kornia.create_meshgrid`` = the five lines of kornia/utils/grid.py it is).  The reference code itself runs unmodified.
The field is the fitted room scene of tests/golden/room_field.npz (a trained field: the recursion is well conditioned on it).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("MNRF_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass


def _create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    """kornia.utils.create_meshgrid (kornia/utils/grid.py): (1,H,W,2) grid of (x, y) pixel coordinates."""
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
    return base.permute(1, 0, 2).unsqueeze(0)


def import_reference_callers():
    """-> (train module, eval module, utils module, datasets.ray_utils module) of the unmodified reference."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _stub("torch_optimizer")
    _stub("pytorch_lightning", LightningModule=type("LightningModule", (object,), {}), Trainer=_Anything)
    _stub("pytorch_lightning.callbacks", ModelCheckpoint=_Anything, TQDMProgressBar=_Anything)
    _stub("pytorch_lightning.loggers", TensorBoardLogger=_Anything)
    _stub("pytorch_lightning.plugins", DDPPlugin=_Anything)
    _stub("kornia", create_meshgrid=_create_meshgrid)
    _stub("kornia.losses", ssim=lambda *a, **k: None)
    _stub("imageio")
    try:
        importlib.import_module("cv2")
    except Exception:
        _stub("cv2")
    train = importlib.import_module("train")
    ev = importlib.import_module("eval")
    utils = importlib.import_module("utils")
    ray_utils = importlib.import_module("datasets.ray_utils")
    return train, ev, utils, ray_utils


def npify(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


# ------------------------------------------------------------------------------------------------ inputs
def room_rays(n, mirror_only=False, seed=0):
    """Rays of the analytic room's camera (mirror_nerf_b200/room_scene.py): a deterministic subset, optionally only pixels
    that look into the mirror (ground truth of the scene)."""
    from mirror_nerf_b200.room_scene import room_pose, trace_room
    from mirror_nerf_b200.synthetic import camera_rays
    allrays = camera_rays(200, 200, c2w=room_pose(1), near=0.05, far=12.0)
    _, m, _ = trace_room(allrays)
    g = np.random.Generator(np.random.PCG64(seed))
    if mirror_only:
        # well inside the mirror (erode the mask by 6 pixels) so that the FITTED mask is 1 there too
        mm = m.view(200, 200)
        inner = mm.clone()
        for s in range(1, 7):
            inner[s:, :] *= mm[:-s, :]; inner[:-s, :] *= mm[s:, :]; inner[:, s:] *= mm[:, :-s]; inner[:, :-s] *= mm[:, s:]
        idx = torch.nonzero(inner.flatten() > 0)[:, 0].numpy()
    else:
        idx = np.arange(allrays.shape[0])
    pick = np.sort(g.choice(idx, size=n, replace=False))
    return allrays[torch.from_numpy(pick)].contiguous(), m[torch.from_numpy(pick)].contiguous()


def room_models(MirrorNeRF):
    from util import room_state_dicts
    sds = room_state_dicts()
    models = {}
    for k, sd in sds.items():
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        models[k] = m
    return models, sds


EVAL_CASES = {
    # name: (n rays, mirror_only, max_recursive_level, roughness?, trace_ray_times, normal_noise_std, seed)
    "l1": (40, False, 1, False, 0, 0.0, 1),
    "l2": (40, False, 2, False, 0, 0.0, 2),
    "rough_l2": (12, True, 2, True, 2, 0.02, 3),
}
TRAIN_CASES = {
    # name: dict(n, gt (use the ground-truth mask) | -1 entry, only_trace_rays_in_mirrors, max level, is_eval, detach flags)
    "gt_only_mirror": dict(n=24, gt=True, only=True, levels=1, is_eval=False, detach_normal=False, detach_ref=False),
    "pred_all_rays_l2": dict(n=16, gt=False, only=False, levels=2, is_eval=True, detach_normal=True, detach_ref=True),
}
TRAIN_SEED = 77


def train_loss(r):
    """Touches every output the reference's losses read (R/losses.py:201-259), both passes."""
    return sum((r[f"rgb_{t}"] ** 2).sum() + r[f"mirror_mask_{t}"].sum() + 0.1 * r[f"normal_dif_{t}"].sum()
               + 0.01 * r[f"depth_{t}"].sum() for t in ("coarse", "fine"))


def write_pl_checkpoint(sds, path=None):
    """A pytorch-lightning style checkpoint of the two fields (attribute names of R/train.py:55,65) plus an unrelated entry;
    seeded values, so the test re-creates the identical file instead of shipping 5 MB of bytes."""
    ck = {"state_dict": {}, "epoch": 3}
    g = torch.Generator().manual_seed(9)
    for tag in ("coarse", "fine"):
        for k, v in sds[tag].items():
            ck["state_dict"][f"nerf_{tag}.{k}"] = torch.randn(v.shape, generator=g)
    ck["state_dict"]["loss.coef"] = torch.ones(1)
    path = path or os.path.join(os.environ.get("TMPDIR", "/tmp"), f"mnrf_golden_ckpt_{os.getpid()}.pt")
    torch.save(ck, path)
    return path


def generate():
    """Run the reference callers; returns {fixture name: {key: ndarray}} (nothing is written)."""
    torch.set_num_threads(8)
    train, ev, utils, ray_utils = import_reference_callers()
    from models.mirror_nerf import Embedding, MirrorNeRF
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    models, sds = room_models(MirrorNeRF)
    for m in models.values():
        m.eval()
    files = {}

    # ---- eval semantics: batched_inference (R/eval.py:114-740) -----------------------------------------------------
    out = {}
    for name, (n, mirror_only, levels, rough, T_extra, std, seed) in EVAL_CASES.items():
        rays, _ = room_rays(n, mirror_only, seed)
        args = types.SimpleNamespace(
            predict_normal=True, only_one_field=False, only_one_field_fine_epoch=2, near=0.05, max_recursive_level=levels,
            app_reflect_newly_placed_objects=False, app_place_new_mirror=False, app_reflection_substitution=False,
            app_control_mirror_roughness=rough, trace_ray_times=T_extra, root_dir="room")
        ev.dataset = types.SimpleNamespace(white_back=False)   # module-level `dataset` of eval.py's __main__ block
        torch.manual_seed(1000 + seed)
        r = ev.batched_inference(models, emb, rays, 64, 128, False, 32768, args=args, trace_secondary_rays=True,
                                 normal_noise_std=std)
        out[f"{name}/rays"] = rays
        for k, v in r.items():
            out[f"{name}/out/{k}"] = v
    files["recursion_eval"] = npify(out)

    # ---- train semantics: NeRFSystem.render_rays_chunk_recursively (R/train.py:129-348) -----------------------------
    out = {}
    for m in models.values():
        m.train()
    for name, c in TRAIN_CASES.items():
        rays, gt = room_rays(c["n"], False, 10 + len(name))
        gt = gt.clone()
        if not c["gt"]:
            gt[0] = -1.0                                   # "no GT mirror mask" marker (R/datasets/blender.py, train.py:157)
        system = object.__new__(train.NeRFSystem)
        system.hparams = types.SimpleNamespace(
            N_samples=64, use_disp=False, perturb=1.0, noise_std=1.0, N_importance=128, chunk=32768,
            trace_secondary_rays=True, only_one_field=False, only_trace_rays_in_mirrors=c["only"],
            max_recursive_level=c["levels"], for_vis=False, detach_normal_in_reflection=c["detach_normal"],
            detach_ref_color_for_blend=c["detach_ref"], train_geometry_stage_end_epoch=0)
        system.current_epoch = 5
        system.train_geometry_stage = False
        system.models = models
        system.embeddings = emb
        system.train_dataset = types.SimpleNamespace(white_back=False)
        for m in models.values():
            m.zero_grad()
        torch.manual_seed(TRAIN_SEED)
        r = system.render_rays_chunk_recursively(
            rays, torch.ones(rays.shape[0]).bool(), recur_level=0, mirror_mask=gt, is_eval=c["is_eval"],
            detach_density_outside_mirror_for_mask_loss=False, detach_density_for_mask_loss=False,
            detach_density_for_normal_loss=False)
        loss = train_loss(r)
        loss.backward()
        out[f"{name}/rays"] = rays
        out[f"{name}/gt_mask"] = gt
        out[f"{name}/loss"] = loss.detach()
        for k, v in r.items():
            out[f"{name}/out/{k}"] = v
        for tag, m in models.items():
            for k, p in m.named_parameters():
                out[f"{name}/grad/{tag}/{k}"] = p.grad.flatten()[::13] if p.grad.numel() > 4096 else p.grad.clone()
                out[f"{name}/gradnorm/{tag}/{k}"] = p.grad.norm()
    files["recursion_train"] = npify(out)

    # ---- helpers either side of the path: checkpoint import, ray generation -------------------------------------------
    out = {}
    import math
    for tag, (H, W) in {"a": (6, 8), "b": (5, 5)}.items():
        focal = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
        a = 0.4
        c2w = torch.tensor([[math.cos(a), 0.0, math.sin(a), 1.0], [0.0, 1.0, 0.0, -0.5], [-math.sin(a), 0.0, math.cos(a), 2.5]])
        dirs = ray_utils.get_ray_directions(H, W, focal)
        ro, rd = ray_utils.get_rays(dirs, c2w)
        out[f"rays_{tag}/HWf"] = np.array([H, W, focal], np.float64)
        out[f"rays_{tag}/c2w"] = c2w
        out[f"rays_{tag}/directions"] = dirs
        out[f"rays_{tag}/rays_o"] = ro
        out[f"rays_{tag}/rays_d"] = rd
    path = write_pl_checkpoint(sds)
    ext = utils.extract_model_state_dict(path, "nerf_fine", prefixes_to_ignore=["normal_net"])
    out["ckpt/extract_keys"] = np.array(sorted(ext.keys()))
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    m.load_state_dict(sds["coarse"])
    utils.load_ckpt(m, path, "nerf_fine", prefixes_to_ignore=["normal_net"])
    for k, v in m.state_dict().items():
        out[f"ckpt/loaded/{k}"] = v[..., :5] if v.dim() == 2 else v[:5]
    files["helpers"] = npify(out)
    return files


def main():
    for name, d in generate().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name + ".npz", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
