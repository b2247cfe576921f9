"""Shared helpers for the parity tests."""
import numpy as np
import torch

from mirror_nerf_b200.synthetic import scene_state_dicts


def T(x, device="cpu"):
    return torch.from_numpy(np.asarray(x)).to(device)


def make_models(device="cuda", predict_normal=True, predict_mirror_mask=True):
    """Two of OUR MirrorNeRF modules loaded with the synthetic state dicts the golden vectors were made with."""
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    models = {}
    for name, sd in scene_state_dicts(predict_normal, predict_mirror_mask).items():
        m = MirrorNeRF(predict_normal=predict_normal, predict_mirror_mask=predict_mirror_mask)
        m.load_state_dict(sd)
        models[name] = m.to(device).eval()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    return models, emb


# every err_stats() call of a GPU test run is logged here and written to gpurun_out/parity_stats.json by conftest.py, so that the
# thresholds in the tests can be kept at "measured x 2" and bench.py / DESIGN.md can quote the measured distributions
RECORDED = []


def _caller_tag():
    import inspect
    import os
    for fr in inspect.stack()[2:8]:
        if fr.function.startswith("test_"):
            return f"{os.path.basename(fr.filename)}::{fr.function}:{fr.lineno}"
    return ""


def err_stats(got, want, floor=None, name=None):
    """Relative error |got-want| / max(|want|, floor) as (median, p99, max, frac>1e-3).  `floor` defaults to the
    RMS of `want` so that near-zero entries are judged on the tensor's own scale."""
    got = got.detach().double().cpu().flatten()
    want = want.detach().double().cpu().flatten()
    if floor is None:
        floor = max(float(want.pow(2).mean().sqrt()), 1e-12)
    e = (got - want).abs() / want.abs().clamp_min(floor)
    if e.numel() == 0:
        return dict(median=0.0, p99=0.0, max=0.0, frac=0.0)
    e = torch.nan_to_num(e, nan=float("inf"))
    q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.double)) if e.numel() < 1_000_000 else \
        torch.tensor([e.median(), e.kthvalue(int(0.99 * e.numel()))[0]])
    out = dict(median=float(q[0]), p99=float(q[1]), max=float(e.max()), frac=float((e > 1e-3).double().mean()))
    RECORDED.append(dict(where=_caller_tag(), name=name, n=int(e.numel()), **out))
    return out


def fmt_stats(name, s):
    return f"{name:28s} median {s['median']:.2e}  p99 {s['p99']:.2e}  max {s['max']:.2e}  frac>1e-3 {s['frac']:.4f}"


def room_state_dicts():
    """The scene-like field: this repo's training path fitted to the analytic room scene (tools/train_room.py, 6000 steps, 33.7 dB;
    stored as fp16, used as fp32).  {"coarse": state_dict, "fine": state_dict}."""
    import os
    from collections import OrderedDict
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "room_field.npz"))
    out = {"coarse": OrderedDict(), "fine": OrderedDict()}
    for k in z.files:
        tag, name = k.split("/", 1)
        out[tag][name] = torch.from_numpy(z[k].astype(np.float32))
    return out
