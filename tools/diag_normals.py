import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import T, make_models
from mirror_nerf_b200.synthetic import scene_state_dicts
from oracle import mirror_nerf_oracle as O
g = dict(np.load(os.path.join(ROOT, "tests/golden/field.npz")))
models, _ = make_models()
m = models["fine"]; m.return_geo_feat = False
x = torch.cat([T(g["xyz"]), T(g["pe_dir"])], 1).cuda()
p = scene_state_dicts()["fine"]
with torch.no_grad():
    gx = O.analytic_normal_explicit(p, T(g["xyz"]))
for impl in ("tc3", "tc1", "fp32"):
    m.field_impl = impl
    with torch.no_grad():
        o = m(x, compute_normal=True, sigma_only=False)
    n = o["normal"].cpu()
    cos = (n * T(g["grad_normal"])).sum(-1)
    bad = (cos < 0.9999).nonzero().flatten().tolist()
    print(impl, "min cos", float(cos.min()), "n_bad", len(bad))
    for i in bad[:12]:
        print("   row", i, "cos %.5f" % float(cos[i]), "xyz", g["xyz"][i].round(3).tolist(), "|grad| %.3e" % float(gx[i].norm()),
              "got", n[i].numpy().round(4).tolist(), "want", g["grad_normal"][i].round(4).tolist())
