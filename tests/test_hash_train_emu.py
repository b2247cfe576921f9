"""CPU check of the hash-grid field's hand-written backward (training).

`mirror_nerf_b200/csrc/hash_train_math.cuh` holds the per-warp math of `k_hash_bwd` (csrc/train_hash.cu).  It compiles as plain
C++; `tests/emu/hash_train_emu.cpp` runs its phases lane by lane.  Here that emulation is compared with torch autograd over
the oracle restatement of MirrorNeRFTcnn.forward (oracle/hashgrid_oracle.py; R/models/mirror_nerf_tcnn.py:151-259), including
the double backward through the analytic normal (create_graph=True, R/utils/func.py:10-25) and the gradient w.r.t. positions
and directions.  The GPU tests (tests/test_gpu_hashgrid.py) then pin the CUDA kernel to the same oracle end to end.
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hashgrid_oracle as HG  # noqa: E402

KEYS = ("sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight",
        "normal_net.0.weight", "normal_net.1.weight", "is_mirror_net.0.weight", "is_mirror_net.0.bias",
        "is_mirror_net.2.weight", "is_mirror_net.2.bias")


class Meta(C.Structure):
    _fields_ = [("bound", C.c_float), ("scale", C.c_float * 16), ("res", C.c_int * 16), ("offset", C.c_uint32 * 16),
                ("size", C.c_uint32 * 16)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libhash_emu.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "emu", "hash_train_emu.cpp")], check=True)
    lib = C.CDLL(out)
    lib.hash_bwd_emu.restype = C.c_int
    lib.hash_bwd_emu2.restype = C.c_int
    return lib


LAYOUTS = {"32x1": "hash_bwd_emu", "16x2": "hash_bwd_emu2"}  # points per warp tile x lanes per point (hash_train_math{,2}.cuh)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(lib, sd, x, d, DR, mirror_on, flags, bound, layout="32x1"):
    lv, _ = HG.level_table(bound)
    M = Meta()
    M.bound = bound
    for i, (scale, res, off, size) in enumerate(lv):
        M.scale[i], M.res[i], M.offset[i], M.size[i] = scale, res, off, size
    nw = lib.hash_bwd_emu_nw()
    wref = np.zeros(nw, np.float32)
    pos = 0
    shapes = HG.param_shapes(bound)
    for k in KEYS:
        n = int(np.prod(shapes[k]))
        if k in sd:
            wref[pos:pos + n] = sd[k].detach().numpy().reshape(-1)
        pos += n
    assert pos == nw
    table = np.ascontiguousarray(sd["encoder.params"].detach().numpy())
    gtable = np.zeros_like(table)
    gsmall = np.zeros(nw, np.float32)
    P = x.shape[0]
    dxd = np.zeros((P, 8), np.float32)
    xs, ds, drs = (np.ascontiguousarray(t, np.float32) for t in (x, d, DR))
    mo = np.ascontiguousarray(mirror_on, np.int32)
    rc = getattr(lib, LAYOUTS[layout])(_fp(table), _fp(wref), C.byref(M), _fp(xs), _fp(ds), _fp(drs), _fp(mo), P,
                          int("normal_net.0.weight" in sd), int("is_mirror_net.0.weight" in sd), int(flags["compute_normal"]),
                          int(flags["detach_normal"]), int(flags["detach_mask"]), 1, int(flags["compute_normal"]),
                          _fp(gtable), _fp(gsmall), _fp(dxd))
    assert rc == 0
    grads, pos = {"encoder.params": gtable}, 0
    for k in KEYS:
        n = int(np.prod(shapes[k]))
        grads[k] = gsmall[pos:pos + n].reshape(shapes[k])
        pos += n
    return grads, dxd


def _oracle(sd, x, d, DR, mirror_on, flags, bound):
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    dt = torch.from_numpy(d).clone().requires_grad_(True)
    mm = torch.from_numpy(mirror_on.astype(np.float32))
    o = HG.field_forward(p, torch.cat([xt, dt], 1), bound=bound, compute_normal=flags["compute_normal"], mirror_mask=mm,
                         detach_density_outside_mirror_for_mask_loss=flags["outside"],
                         detach_density_for_mask_loss=flags["detach_mask"],
                         detach_density_for_normal_loss=flags["detach_normal"])
    g = torch.from_numpy(DR)
    loss = (o["sigma"][:, 0] * g[:, 0]).sum() + (o["rgb"] * g[:, 1:4]).sum()
    if "is_mirror" in o:
        loss = loss + (o["is_mirror"][:, 0] * g[:, 4]).sum()
    if "pred_normal" in o:
        loss = loss + (o["pred_normal"] * g[:, 5:8]).sum()
    if flags["compute_normal"]:
        loss = loss + (o["normal"] * g[:, 8:11]).sum()
    loss.backward()
    return {k: v.grad.numpy() for k, v in p.items()}, xt.grad.numpy(), dt.grad.numpy()


CASES = {
    "full": dict(compute_normal=True, detach_normal=False, detach_mask=False, outside=False),
    "no_analytic_normal": dict(compute_normal=False, detach_normal=False, detach_mask=False, outside=False),
    "detach_normal": dict(compute_normal=True, detach_normal=True, detach_mask=False, outside=False),
    "detach_mask": dict(compute_normal=True, detach_normal=False, detach_mask=True, outside=False),
    "detach_outside_mirror": dict(compute_normal=True, detach_normal=False, detach_mask=False, outside=True),
    "no_heads": dict(compute_normal=True, detach_normal=False, detach_mask=False, outside=False, heads=False),
}


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("case", list(CASES))
def test_hash_backward_math_matches_autograd(emu, case, layout):
    flags = dict(CASES[case])
    heads = flags.pop("heads", True)
    bound = 1.0
    sd = HG.make_state_dict(seed=3, bound=bound, sigma_scale=4.0, predict_normal=heads, predict_mirror_mask=heads)
    rng = np.random.default_rng(11)
    P = 77  # two full warps and a ragged tail
    x = rng.uniform(-0.98, 0.98, size=(P, 3)).astype(np.float32)
    d = rng.normal(size=(P, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    DR = rng.normal(size=(P, 12)).astype(np.float32)
    DR[:, 11] = 0
    if not flags["compute_normal"]:
        DR[:, 8:11] = 0
    if not heads:
        DR[:, 4:8] = 0
    mirror_on = np.ones(P, np.int32)
    if flags["outside"]:
        mirror_on = (rng.uniform(size=P) < 0.5).astype(np.int32)
    got, dxd = _run(emu, sd, x, d, DR, mirror_on, flags, bound, layout)
    want, gx, gd = _oracle(sd, x, d, DR, mirror_on, flags, bound)
    for k, w in want.items():
        g = got[k].reshape(w.shape)
        scale = np.abs(w).max()
        assert scale > 0, k
        err = np.abs(g - w).max() / scale
        assert err < 1e-5, (k, err)  # fp32 accumulation order only (measured <= 7e-7)
    sx = np.abs(gx).max()
    assert np.abs(dxd[:, 0:3] - gx).max() / sx < 1e-5
    sd_ = np.abs(gd).max()
    assert np.abs(dxd[:, 3:6] - gd).max() / sd_ < 1e-5


def _setup(P, seed, heads=True):
    sd = HG.make_state_dict(seed=5, bound=1.0, sigma_scale=4.0, predict_normal=heads, predict_mirror_mask=heads)
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.98, 0.98, size=(P, 3)).astype(np.float32)
    d = rng.normal(size=(P, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    DR = rng.normal(size=(P, 12)).astype(np.float32)
    DR[:, 11] = 0
    return sd, x, d, DR


def test_hash_backward_is_linear_in_the_output_gradients(emu):
    """Size-independent property: for fixed points the backward is a linear map of the per-point gradient record, second-order
    path included (the analytic normal's double backward is linear in d L / d n)."""
    flags = dict(CASES["full"])
    sd, x, d, DR1 = _setup(64, 21)
    DR2 = np.random.default_rng(22).normal(size=DR1.shape).astype(np.float32)
    DR2[:, 11] = 0
    on = np.ones(64, np.int32)
    g1, q1 = _run(emu, sd, x, d, DR1, on, flags, 1.0)
    g2, q2 = _run(emu, sd, x, d, DR2, on, flags, 1.0)
    g3, q3 = _run(emu, sd, x, d, 2.0 * DR1 - 0.5 * DR2, on, flags, 1.0)
    for k in g1:
        want = 2.0 * g1[k] - 0.5 * g2[k]
        scale = max(np.abs(want).max(), 1e-20)
        assert np.abs(g3[k] - want).max() / scale < 2e-5, k
    want = 2.0 * q1 - 0.5 * q2
    assert np.abs(q3 - want).max() / np.abs(want).max() < 2e-5


def test_hash_backward_is_additive_over_point_shards(emu):
    """Parameter gradients of a batch = sum over any split of its points (tiles, ragged tails and the order of the scatter-adds
    do not matter beyond fp32 rounding): the property data-parallel training relies on."""
    flags = dict(CASES["full"])
    sd, x, d, DR = _setup(101, 23)
    on = np.ones(101, np.int32)
    whole, dxd = _run(emu, sd, x, d, DR, on, flags, 1.0)
    parts = [(0, 37), (37, 38), (38, 101)]
    acc, rows = None, []
    for lo, hi in parts:
        g, q = _run(emu, sd, x[lo:hi], d[lo:hi], DR[lo:hi], on[lo:hi], flags, 1.0)
        rows.append(q)
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    for k in whole:
        scale = max(np.abs(whole[k]).max(), 1e-20)
        assert np.abs(acc[k] - whole[k]).max() / scale < 2e-5, k
    assert np.array_equal(np.concatenate(rows, 0), dxd)  # per-point ray-gradient records do not depend on the tiling


def test_hash_backward_zero_gradient_and_empty_heads(emu):
    """Edge cases: an all-zero gradient record gives exactly zero everywhere; a model without heads leaves the head slots at zero."""
    flags = dict(CASES["full"])
    sd, x, d, DR = _setup(40, 25)
    g, q = _run(emu, sd, x, d, np.zeros_like(DR), np.ones(40, np.int32), flags, 1.0)
    assert all(float(np.abs(v).max()) == 0.0 for v in g.values()) and float(np.abs(q).max()) == 0.0
    sd2, x, d, DR = _setup(40, 26, heads=False)
    DR[:, 4:8] = 0
    g, _ = _run(emu, sd2, x, d, DR, np.ones(40, np.int32), flags, 1.0)
    for k in ("normal_net.0.weight", "normal_net.1.weight", "is_mirror_net.0.weight", "is_mirror_net.2.bias"):
        assert float(np.abs(g[k]).max()) == 0.0, k
    assert float(np.abs(g["encoder.params"]).max()) > 0


@pytest.mark.parametrize("layout", list(LAYOUTS))
def test_hash_backward_steps_do_not_depend_on_lane_order(emu, layout):
    """On the GPU the lanes of a step run concurrently; here they run one after the other.  If some step read, in one lane, a
    buffer element another lane writes in the same step (a missing barrier), the two lane orders would disagree."""
    flags = dict(CASES["full"])
    sd, x, d, DR = _setup(53, 31)
    on = np.ones(53, np.int32)
    try:
        emu.hash_bwd_emu_set_reverse(0)
        g0, q0 = _run(emu, sd, x, d, DR, on, flags, 1.0, layout)
        emu.hash_bwd_emu_set_reverse(1)
        g1, q1 = _run(emu, sd, x, d, DR, on, flags, 1.0, layout)
    finally:
        emu.hash_bwd_emu_set_reverse(0)
    assert np.array_equal(q0, q1)
    for k in g0:
        # the accumulation order of the sums over points differs between the two runs: compare to rounding, not bit for bit
        scale = max(np.abs(g0[k]).max(), 1e-20)
        assert np.abs(g0[k] - g1[k]).max() / scale < 1e-5, k


def test_emulation_runs_the_same_step_sequence_as_the_kernels():
    """The emulation harness repeats the kernels' step sequence by hand; keep the two in lockstep (order and names of the
    phase_* / step_* calls of k_hash_bwd / k_hash_bwd2 vs hash_bwd_emu / hash_bwd_emu2)."""
    import re
    cu = open(os.path.join(ROOT, "mirror_nerf_b200", "csrc", "train_hash.cu")).read()
    emu_src = open(os.path.join(ROOT, "tests", "emu", "hash_train_emu.cpp")).read()

    def calls(text, prefix):
        return re.findall(r"\b(" + prefix + r"_[a-z0-9]+)\(", text)

    k1 = cu[cu.index("k_hash_bwd("):cu.index("k_hash_bwd2(")]
    k2 = cu[cu.index("k_hash_bwd2("):cu.index("int hash_bwd_layout()")]
    e1 = emu_src[emu_src.index('int hash_bwd_emu('):emu_src.index('int hash_bwd_emu_nw')]
    e2 = emu_src[emu_src.index('int hash_bwd_emu2('):]
    assert calls(k1, "phase") == calls(e1, "phase") and len(calls(k1, "phase")) == 14
    assert calls(k2, "step") == calls(e2, "step") and len(calls(k2, "step")) == 26
    # every step of the two-lanes-per-point kernel is followed by a warp barrier before the next one (step_n / step_o touch
    # disjoint rows and share one)
    body = k2[k2.index("step_a1("):k2.index("if (dxd != nullptr")]
    stmts = [s.strip() for s in re.split(r";|\{|\}", body) if s.strip()]
    names = [re.match(r"(?:if \(so\)\s*)?(step_[a-z0-9]+|__syncwarp)\(", s) for s in stmts]
    seq = [m.group(1) for m in names if m]
    for a, b in zip(seq, seq[1:]):
        if a.startswith("step_") and b.startswith("step_"):
            assert (a, b) == ("step_n", "step_o"), (a, b)
