"""Drop-in installer: make the reference's own scripts call this renderer without editing them.

R/train.py:14 does ``from models.rendering import *`` and R/eval.py:10 ``from models.rendering import render_rays``; both bind
whatever ``models.rendering`` exposes at import time.  ``install()`` imports the reference's ``models.rendering`` (it needs only
torch + einops) and replaces its ``render_rays`` / ``sample_pdf`` attributes with ours, so a launcher is three lines::

    import mirror_nerf_b200.dropin as dropin
    dropin.install("/path/to/Mirror-NeRF")
    runpy.run_path("/path/to/Mirror-NeRF/eval.py", run_name="__main__")     # or train.py

The reference's model classes (R/models/mirror_nerf.py, mirror_nerf_tcnn.py) can stay as they are: ``render_rays`` only reads
their parameters (state_dict keys) and the embeddings' ``N_freqs``.
"""
from __future__ import annotations

import importlib
import sys
import types


def install(reference_root: str | None = None):
    """Patch the reference's ``models.rendering`` in ``sys.modules``.  Returns the patched module."""
    from . import rendering as fast
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # R/utils/__init__.py:7 imports torch_optimizer at module import; the hot path never uses it
    if "torch_optimizer" not in sys.modules:
        try:
            importlib.import_module("torch_optimizer")
        except ImportError:
            sys.modules["torch_optimizer"] = types.ModuleType("torch_optimizer")
    ref = importlib.import_module("models.rendering")
    if not hasattr(ref, "render_rays"):
        raise ImportError("models.rendering does not look like Mirror-NeRF's (no render_rays)")
    ref._reference_render_rays = getattr(ref, "_reference_render_rays", ref.render_rays)
    ref._reference_sample_pdf = getattr(ref, "_reference_sample_pdf", ref.sample_pdf)
    ref.render_rays = fast.render_rays
    ref.sample_pdf = fast.sample_pdf
    return ref


def uninstall():
    ref = sys.modules.get("models.rendering")
    if ref is not None and hasattr(ref, "_reference_render_rays"):
        ref.render_rays = ref._reference_render_rays
        ref.sample_pdf = ref._reference_sample_pdf
