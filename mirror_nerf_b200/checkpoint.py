"""Weight import: restates R/utils/__init__.py:109-136.  A PyTorch-Lightning checkpoint of the
reference stores both fields under the prefixes ``nerf_coarse.`` / ``nerf_fine.`` (R/train.py:55,65); a plain ``state_dict``
file works too.  The tensors keep the reference layout ([out,in] fp32), which is exactly what ``packed_field`` ingests."""
from __future__ import annotations

import torch


def extract_model_state_dict(ckpt, model_name="model", prefixes_to_ignore=()):
    """`ckpt`: path or already-loaded dict.  Returns {key without '<model_name>.': tensor}."""
    checkpoint = torch.load(ckpt, map_location="cpu") if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__") else ckpt
    if "state_dict" in checkpoint:  # pytorch-lightning checkpoint
        checkpoint = checkpoint["state_dict"]
    out = {}
    for k, v in checkpoint.items():
        if not k.startswith(model_name):
            continue
        k = k[len(model_name) + 1:]
        if any(k.startswith(p) for p in prefixes_to_ignore):
            continue
        out[k] = v
    return out


def load_ckpt(model, ckpt, model_name="model", prefixes_to_ignore=()):
    """Same contract as the reference's load_ckpt: silently returns on an empty path, asserts that the model is present."""
    if not ckpt:
        return
    model_dict = model.state_dict()
    found = extract_model_state_dict(ckpt, model_name, prefixes_to_ignore)
    assert len(found) > 0, f"[Error] can not find {model_name} in checkpoint"
    model_dict.update(found)
    model.load_state_dict(model_dict, strict=False)
