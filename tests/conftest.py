import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    """Measured error distributions of this run -> gpurun_out/parity_stats.json (GPU runs only)."""
    try:
        import json
        import torch
        import util
        if not util.RECORDED or not torch.cuda.is_available():
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_stats.json"), "w") as f:
            json.dump(util.RECORDED, f, indent=0)
    except Exception:
        pass


REFERENCE = os.environ.get("MNRF_REFERENCE", "/root/reference")


@pytest.fixture(scope="session", params=["file", "live"])
def golden(request):
    """Reference outputs on the seeded inputs.  "file": the committed tests/golden/*.npz (made on another host CPU, so
    float comparisons are host-tolerant).  "live": the unmodified reference imported from /root/reference and run in
    this process (build container only; skipped elsewhere) -- same ATen kernels on the same host, so the oracle must
    reproduce it exactly."""
    import numpy as np

    if request.param == "file":
        def load(name):
            return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        load.exact = False
        return load
    if not os.path.isdir(os.path.join(REFERENCE, "models")):
        pytest.skip("reference tree not present on this machine")
    sys.path.insert(0, GOLDEN)
    files = {}

    def load(name):
        if name not in files:
            # two generators: the render_rays level (make_golden.py) and its callers / helpers (make_golden_recursion.py)
            import make_golden
            import make_golden_recursion
            gen = make_golden_recursion if name in ("recursion_eval", "recursion_train", "helpers") else make_golden
            files.update(gen.generate())
        return dict(files[name])
    load.exact = True
    return load
