"""CPU-only checks (no compute calls): the C-ABI library loads and exports every symbol include/mnrf.h declares,
the ctypes structs match the header, the host layer refuses non-CUDA inputs loudly, and the module mirrors the
reference's state_dict layout."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mnrf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnrf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mnrf.h but not exported by libmnrf.so"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.mnrf_abi_version() == 2
    assert lib.mnrf_macs_full() == 659456 and lib.mnrf_macs_sigma_only() == 524416


def test_struct_sizes_match_header():
    from mirror_nerf_b200 import _lib
    assert C.sizeof(_lib.CompositeOut) == 10 * 8
    assert C.sizeof(_lib.LevelCfg) == 10 * 4 + 4 + 4 + 8 + 8      # ABI 2: + early_termination_eps, no_fused_composite, 2 pointers
    assert C.sizeof(_lib.TraceCfg) == 64 + 3 * 4 + 4 + 8
    assert C.sizeof(_lib.TraceOut) == 11 * 8
    assert C.sizeof(_lib.LevelRng) == 4 * 8
    assert C.sizeof(_lib.LevelOut) == 8 + 80 + 8 + 8 + 80 + 8
    lib = _lib.load()
    cfg = _lib.LevelCfg(n_samples=64, n_importance=128)
    # dirbias n*128*4 + coarse n*64*32 + fine n*192*32 bytes + the fused compositor's work counter
    assert lib.mnrf_level_workspace_bytes(1000, C.byref(cfg)) == 1000 * (512 + 2048 + 6144) + 256


def test_compute_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mirror_nerf_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    arr = (C.c_void_p * 32)(*[1] * 32)
    assert lib.mnrf_field_create(C.byref(h), arr, None) != 0
    assert b"cuda" in lib.mnrf_last_error().lower()


def test_host_layer_rejects_cpu_tensors():
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    from mirror_nerf_b200.rendering import render_rays, sample_pdf
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    with pytest.raises(RuntimeError, match="CUDA"):
        render_rays({"coarse": m}, emb, torch.zeros(4, 8), 64, False, 0, 0, 0, 1024, False)
    with pytest.raises(RuntimeError, match="rays must be"):
        render_rays({"coarse": m}, emb, torch.zeros(4, 7), 64, False, 0, 0, 0, 1024, False)
    with pytest.raises(RuntimeError, match="CUDA"):
        sample_pdf(torch.zeros(2, 63), torch.zeros(2, 62), 8, det=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(3, 30))
    with pytest.raises(RuntimeError, match="CUDA"):
        emb["xyz"](torch.zeros(3, 3))


def test_module_mirrors_reference_state_dict():
    from mirror_nerf_b200.mirror_nerf import PARAM_KEYS, MirrorNeRF
    from mirror_nerf_b200.synthetic import field_param_shapes, make_state_dict
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    sd = m.state_dict()
    shapes = field_param_shapes()
    assert list(sd) == list(shapes)  # same keys, same order as the reference module
    assert set(sd) == set(PARAM_KEYS)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    m.load_state_dict(make_state_dict(3))
    assert sum(p.numel() for p in m.parameters()) == 662152
    m0 = MirrorNeRF()
    assert sum(p.numel() for p in m0.parameters()) == 595844
    with pytest.raises(NotImplementedError):
        MirrorNeRF(D=4)


def test_shard_plan():
    from mirror_nerf_b200.parallel import shard_bounds
    n = 640000
    for world in (1, 2, 4, 8, 3):
        b = [shard_bounds(n, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 128  # whole 128-ray tiles per rank
    assert shard_bounds(5, 3, 8) == (5, 5)  # more ranks than tiles: empty shards are legal


def test_checkpoint_import_pl_prefixes(tmp_path):
    """R/utils/__init__.py:109-136: a Lightning checkpoint keeps both fields under nerf_coarse./nerf_fine. prefixes."""
    from mirror_nerf_b200.checkpoint import extract_model_state_dict, load_ckpt
    from mirror_nerf_b200.mirror_nerf import MirrorNeRF
    from mirror_nerf_b200.synthetic import make_state_dict
    sd_c, sd_f = make_state_dict(7), make_state_dict(8)
    pl = {"epoch": 3, "state_dict": {**{f"nerf_coarse.{k}": v for k, v in sd_c.items()},
                                     **{f"nerf_fine.{k}": v for k, v in sd_f.items()},
                                     "nerf_fine.normal_net_bg.0.weight": torch.zeros(1)}}
    path = tmp_path / "epoch=3.ckpt"
    torch.save(pl, path)
    got = extract_model_state_dict(str(path), "nerf_fine", prefixes_to_ignore=["normal_net_bg"])
    assert set(got) == set(sd_f)
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    load_ckpt(m, str(path), "nerf_coarse")
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd_c[k]), k
    load_ckpt(m, "", "nerf_coarse")  # empty path: silently ignored like the reference
    with pytest.raises(AssertionError):
        load_ckpt(m, str(path), "nerf_missing")


def test_hashgrid_module_mirrors_reference_state_dict_and_level_table():
    """nerf_tcnn model family: state_dict keys/shapes of R/models/mirror_nerf_tcnn.py, tinycudann's parameter count for the
    configuration (16 levels x 2 features, 2^19, base 16, finest 2048 -> 12,196,240), and the host level table == the oracle's."""
    from mirror_nerf_b200.mirror_nerf_tcnn import MirrorNeRFTcnn, level_table, n_encoder_params
    from oracle import hashgrid_oracle as H
    assert n_encoder_params(1.0) == 12196240 == H.n_encoder_params(1.0)
    for bound in (1.0, 2.0, 0.5):
        assert level_table(bound) == H.level_table(bound)
    m = MirrorNeRFTcnn(bound=1, predict_normal=True, predict_mirror_mask=True)
    sd = m.state_dict()
    shapes = H.param_shapes()
    assert list(sd) == list(shapes)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    with pytest.raises(NotImplementedError):
        MirrorNeRFTcnn(hidden_dim=128)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 6), compute_normal=False)


def test_hashgrid_oracle_properties():
    """The restated encoder: C0-continuous across cell boundaries, level-major layout, dense levels index the table directly."""
    from oracle import hashgrid_oracle as H
    lv, total = H.level_table(1.0)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(total * 2, generator=g)
    x = torch.rand(64, 3, generator=g)
    e0 = H.hashgrid_encode(table, x)
    assert e0.shape == (64, 32)
    e1 = H.hashgrid_encode(table, x + 1e-6)
    assert float((e0 - e1).abs().max()) < 5e-2  # no jumps: |d enc| <= scale_max * |dx| * |table| ~ 2048 * 1e-6 * 4
    # level 0 is dense (16^3 = 4096 entries): a point at the centre-offset grid node reads exactly one entry
    scale, res, off, size = lv[0]
    node = torch.tensor([[3, 5, 7]], dtype=torch.float32)
    xn = (node - 0.5) / scale
    e = H.hashgrid_encode(table, xn)
    idx = 3 + 5 * res + 7 * res * res
    assert torch.allclose(e[0, :2], table.view(-1, 2)[off + idx], atol=1e-5)
    assert H.sh4(torch.tensor([[0.0, 0.0, 1.0]])).shape == (1, 16)


def test_dropin_install_patches_reference_module():
    """after install(), `from models.rendering import render_rays` (R/eval.py:10, R/train.py:14) binds
    ours; the signature keeps the reference's positional parameters (R/models/rendering.py:54-67)."""
    import inspect
    import sys
    ref_root = os.environ.get("MNRF_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "models")):
        pytest.skip("reference tree not present on this machine")
    from mirror_nerf_b200 import dropin, rendering
    mod = dropin.install(ref_root)
    try:
        ns = {}
        exec("from models.rendering import render_rays, sample_pdf", ns)
        assert ns["render_rays"] is rendering.render_rays and ns["sample_pdf"] is rendering.sample_pdf
        ours = list(inspect.signature(rendering.render_rays).parameters)
        theirs = list(inspect.signature(mod._reference_render_rays).parameters)
        assert ours[:len(theirs) - 1] == theirs[:-1] and ours[-1] == theirs[-1] == "kwargs", (ours, theirs)
        for name in theirs[:-1]:
            a = inspect.signature(rendering.render_rays).parameters[name].default
            b = inspect.signature(mod._reference_render_rays).parameters[name].default
            assert a == b, (name, a, b)
    finally:
        dropin.uninstall()
    assert sys.modules["models.rendering"].render_rays is mod._reference_render_rays


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) works without a GPU and prints exactly one JSON
    line on stdout carrying the contract's keys; ours must refuse to run without a GPU instead of falling back."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-rays", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                           text=True, timeout=300)
        assert r.returncode != 0 and not r.stdout.strip(), "ours must fail loudly without a GPU, not fall back to the CPU"


def test_traffic_record_is_reproducible_from_the_committed_launch_list(tmp_path):
    """bench.py reports `roofline.traffic` from profiles/r02_field_tc_traffic.json; that record must be what
    tools/ncu_launch_summary.py derives from the committed ncu launch list of the bench command (no hand-edited constant)."""
    import gzip
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = tmp_path / "launches.csv"
    with gzip.open(os.path.join(root, "profiles", "r02_v5_bench_launches.csv.gz"), "rb") as f:
        csv_path.write_bytes(f.read())
    out_txt, out_json = tmp_path / "summary.txt", tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_launch_summary.py"), str(csv_path), str(out_txt), str(out_json)],
                   check=True, capture_output=True)
    got = json.loads(out_json.read_text())
    want = json.load(open(os.path.join(root, "profiles", "r02_field_tc_traffic.json")))
    for k in ("launches", "rays_per_launch", "dram_bytes_per_launch_avg", "dram_bytes_per_ray"):
        assert got[k] == want[k], (k, got[k], want[k])
    # the fused fine pass moves ~1.4 KB per ray through DRAM (the unfused sequence of round 1: 17.6 KB)
    assert 1000 < want["dram_bytes_per_ray"] < 2000
