#!/usr/bin/env python
"""bench.py -- rays/sec of the render_rays hot path on B200 (BASELINE.json metric).

Workload (config[1] of BASELINE.json): synthetic 800x800 view (640,000 primary rays), 64 coarse + 128 importance
samples (192 fine points), both heads, eval mode, 1 reflection bounce with the reference's eval semantics (level 0,
then every ray of the batch is reflected and re-rendered once, then blended by the thresholded mirror mask:
R/eval.py:132-160,545-548,676-697) -> 2 render levels per primary ray.  A "step" is one such image.  Multi-GPU: one
process per GPU, each rank renders its own view (weak scaling), no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--field-impl tc3|tc2|tc1]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.  Both arms print the same
`config` object (the workload); what is specific to an arm's run sits under `run`.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 800
N_SAMPLES, N_IMPORTANCE = 64, 128
FLOP_PER_RAY_LEVEL = 320.36e6  # 64*S + 192*F MACs * 2 (unpadded reference layer sizes)
LEVELS = 2                     # level 0 + 1 bounce (eval semantics re-traces all rays of the batch)
METRIC = "rays/sec (64c+128f samples, 1 bounce)"
WORKLOAD = "synthetic 800x800 mirror scene, 64+128 samples, 1 reflection bounce, eval semantics"
# the workload both arms are measured on (BASELINE.json configs[1]); identical in the two JSON lines
CONFIG = {"workload": WORKLOAD, "image": "800x800 synthetic pinhole view (R/datasets/ray_utils.py:6-53), near 0.05, far 8.0",
          "samples": "64 coarse + 128 importance (192 fine points per ray)", "bounces": 1, "levels_per_ray": LEVELS,
          "semantics": "eval: R/eval.py::batched_inference, perturb = noise_std = 0, test_time, predicted normals + mirror mask",
          "field": "two MirrorNeRF (D=8, W=256, skip 4, 10/4 frequencies, normal + mirror heads), synthetic weights "
                   "(mirror_nerf_b200/synthetic.py: sigma head x40, mirror head x100)"}


def view_pose(i):
    """Small orbit of camera poses (one per rank / step): rotate about y, camera 2.5 away looking at the origin."""
    import math
    import torch
    a = 0.35 * i
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0.0, s, 2.5 * s], [0.0, 1.0, 0.0, 0.0], [-s, 0.0, c, 2.5 * c]])


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.power, self.power_limit = [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                self.power_limit = None
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s, pw = sorted(self.samples), sorted(self.power)
        # board power under load against the enforced limit: with sw_power_cap active the SM clock is whatever the power budget
        # allows, i.e. throughput follows energy per ray, not cycles per ray
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w": (pw[len(pw) // 2] if pw else None), "power_limit_w": self.power_limit}


def cpu_reference_rate(n_rays, steps, warmup, threads):
    """The reference algorithm (oracle restatement = the same ATen CPU kernels the reference runs) on the host cores:
    `n_rays` rays of the same view per step, 64+128 samples, 1 bounce, eval semantics.  Returns (rays/s, rgb, rays)."""
    import torch
    from mirror_nerf_b200.synthetic import camera_rays, scene_state_dicts
    from oracle import mirror_nerf_oracle as O
    torch.set_num_threads(threads)
    params = scene_state_dicts()
    allrays = camera_rays(H, W, c2w=view_pose(0))
    idx = torch.linspace(0, allrays.shape[0] - 1, n_rays).long()
    rays = allrays[idx].contiguous()
    fn = lambda r: O.render_rays(params, r, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False, test_time=True,
                                 compute_normal=False)
    out = None
    per_step = []
    with torch.no_grad():
        for _ in range(warmup):
            O.trace_eval(fn, rays[:256], 1)   # warm-up on 256 rays
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            out = O.trace_eval(fn, rays, 1)
            per_step.append(time.perf_counter() - t1)
        dt = time.perf_counter() - t0
    cpu_reference_rate.best = n_rays / min(per_step)
    return n_rays * steps / dt, out["rgb_fine"], rays, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_rays
    rate, _, _, step_s = cpu_reference_rate(n, args.steps, args.warmup, threads)
    sample = (f"{n} rays of the 800x800 view per step (BASELINE.md section 4: N = 4096, warm-up 256 rays), both levels of "
              f"the bounce, {args.steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG,
            "run": {"sample": sample, "threads": threads, "best_step_rays_per_s": cpu_reference_rate.best,
                    "implementation": "oracle port of the reference's torch CPU path (same ATen kernels; /root/reference cannot "
                                      "travel to the GPU box), pinned bit-exactly to the reference in tests/test_oracle_*.py"},
            "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


TRAIN_RAYS = 4096             # BASELINE config 5: 4096-ray batches (per process, PL semantics: R/train.py:368-375)
TRAIN_FLOP_PER_RAY = 1.77e9   # 256 points x ~6.9 MFLOP (forward + analytic-normal chain + backward + double backward)


def train_loss(r, target, mask_gt):
    """Stand-in for R/losses.py:201-259 (colour MSE on both passes, mirror-mask BCE, normal consistency 1e-4,
    normal regularisation 1e-4): tiny per-ray torch ops, every differentiable output of the level is used."""
    import torch
    loss = 0.0
    for t in ("coarse", "fine"):
        loss = loss + ((r[f"rgb_{t}"] - target) ** 2).mean()
        m = r[f"mirror_mask_{t}"].clamp(1e-7, 1 - 1e-7)
        loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(m, mask_gt)
        loss = loss + 1e-4 * r[f"normal_dif_{t}"].mean()
        loss = loss + 1e-4 * (torch.relu(r[f"pred_normal_{t}"] * r["_rays_d"].unsqueeze(1)).sum(-1) * r[f"weights_{t}"]).mean()
    return loss


def train_gemm_rooflines(dev, hbm_peak, tensor_peak):
    """The two layer GEMMs that make up 90 % of a training step (csrc/train_tc.cu), each timed alone on the operands of one
    4096-ray batch (786,432 points x 256 features): the bias + ReLU flavour of the forward trunk (k_gemm_tc_nn<1>) and the weight
    gradient (k_gemm_tc_tn).  Both bounds are reported: HBM (every fp32 activation is read once and written once per layer:
    algorithmic bytes) and tensor (three tf32 passes per MAC; kind::tf32 runs at half the 16-bit rate)."""
    import ctypes as C
    import torch
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.mirror_nerf import MirrorNeRF, packed_field
    from mirror_nerf_b200.synthetic import make_state_dict
    lib = _lib.load()
    m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
    m.load_state_dict(make_state_dict(1))
    pf = packed_field(m.to(dev))
    P = TRAIN_RAYS * (N_SAMPLES + N_IMPORTANCE)
    out = {}
    for kind, step, name, what in ((0, 1, "k_gemm_tc_nn", "C[P,256] = relu(A[P,256] W^T + b): reads A, writes C"),
                                   (1, 0, "k_gemm_tc_tn", "dW[256,256] += A[P,256]^T B[P,256]: reads A and B")):
        ms = C.c_float()
        _lib.check(lib.mnrf_debug_gemm_bench(pf.handle, kind, step, P, 1, 0, 10, C.byref(ms)), "mnrf_debug_gemm_bench")
        nbytes = 2.0 * P * 256 * 4
        gbs = nbytes / (ms.value * 1e-3) / 1e9
        issued = 3 * 2.0 * P * 256 * 256 / (ms.value * 1e-3) / 1e12   # three tf32 passes
        out[name] = {"what": what, "points": P, "ms_per_launch": ms.value, "bytes_per_launch": nbytes,
                     "hbm_GBps": gbs, "hbm_frac": gbs / hbm_peak,
                     "tf32_tflops_issued": issued, "tensor_frac": issued / (tensor_peak / 2.0)}
    torch.cuda.empty_cache()
    return out


def train_bench(dev, world, rank, steps, warmup, peer_fused=False):
    """BASELINE config 5 on this rank: 4096-ray batch, train semantics (perturb=1, noise_std=1, analytic normals), forward +
    backward through csrc/train.cu + train_tc.cu, ONE flat NCCL all-reduce of the 5.3 MB gradient buffer, one Adam kernel.  Returns a dict."""
    import torch
    import torch.distributed as dist
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    from mirror_nerf_b200.parallel import FlatDataParallel
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import camera_rays, scene_state_dicts
    models = {}
    for k, sd in scene_state_dicts().items():
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        models[k] = m.to(dev).train()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    ddp = FlatDataParallel(models, lr=5e-4, peer_fused=peer_fused and world > 1)
    g = torch.Generator().manual_seed(1234 + rank)
    allrays = camera_rays(H, W, c2w=view_pose(rank))
    rays = allrays[torch.randperm(allrays.shape[0], generator=g)[:TRAIN_RAYS]].contiguous().to(dev)
    target = torch.rand(TRAIN_RAYS, 3, generator=g).to(dev)
    mask_gt = (torch.rand(TRAIN_RAYS, generator=g) > 0.7).float().to(dev)
    losses = []

    def one():
        ddp.zero_grad()
        r = render_rays(models, emb, rays, N_SAMPLES, False, 1.0, 1.0, N_IMPORTANCE, 32768, False, test_time=False,
                        compute_normal=True)
        r["_rays_d"] = rays[:, 3:6]
        loss = train_loss(r, target, mask_gt)
        loss.backward()
        ddp.step()
        losses.append(loss.detach())

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    rate = world * TRAIN_RAYS * steps / (ms * 1e-3)
    return {"metric": "rays/sec (train step: forward + backward + gradient all-reduce + Adam; 4096-ray batch per GPU, "
                      "64+128 samples, analytic normals, fp32)",
            "value": rate, "unit": "rays/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "dtype": "tf32x3-split operands, f32 accumulate (tcgen05 kind::tf32; fp32-grade)", "launches_per_step": (_lib.launch_count() - l0) / steps,
            "algorithmic_tflops": rate / world * TRAIN_FLOP_PER_RAY / 1e12,
            "allreduce_bytes_per_step": ddp.flat_grads.numel() * 4 if world > 1 else 0,
            "exchange": ("none" if world == 1 else ("fused peer-memory reduce-scatter + Adam + all-gather kernel (NVLink P2P)"
                                                    if ddp.peer is not None else "ncclAllReduce (flat buffer) + Adam kernel")),
            "loss_first": float(losses[0]), "loss_last": float(losses[-1])}


def hash_models(dev, sigma_scale=20.0):
    """Synthetic nerf_tcnn model pair: table ~ U(-1,1) (a trained table is O(1)), nn.Linear-style small MLPs."""
    import numpy as np
    import torch
    from mirror_nerf_b200.mirror_nerf_tcnn import MirrorNeRFTcnn
    models = {}
    for k, seed in (("coarse", 7), ("fine", 8)):
        m = MirrorNeRFTcnn(bound=1, predict_normal=True, predict_mirror_mask=True)
        g = np.random.Generator(np.random.PCG64(seed))
        with torch.no_grad():
            for name, p in m.named_parameters():
                lim = 1.0 if name == "encoder.params" else 1.0 / (p.shape[-1] ** 0.5)
                p.copy_(torch.from_numpy(g.uniform(-lim, lim, size=tuple(p.shape)).astype(np.float32)))
            m.sigma_net[1].weight[0] *= sigma_scale
        models[k] = m.to(dev).eval()
    return models


def hash_train_bench(dev, steps, warmup=2):
    """BASELINE config 3 under train.py semantics: one optimisation step of the hash-grid model pair on a 4096-ray batch
    (64+128 samples, perturb=1, noise_std=1, analytic normals with their double backward): forward (csrc/field_hash.cu) +
    backward (csrc/train_hash.cu: recompute, table scatter, small-MLP gradients) + one Adam kernel over the flat 24.4 M-parameter
    buffer (two tables of 12.2 M fp32 entries).  Returns a dict."""
    import torch
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.mirror_nerf import Embedding
    from mirror_nerf_b200.parallel import FlatDataParallel
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import camera_rays
    models = hash_models(dev, sigma_scale=5.0)
    for m in models.values():
        m.train()
    emb = {"xyz": Embedding(0), "dir": Embedding(0)}
    ddp = FlatDataParallel(models, lr=1e-3)
    g = torch.Generator().manual_seed(4321)
    c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.9]])
    allrays = camera_rays(H, W, c2w=c2w, near=0.05, far=2.0)
    rays = allrays[torch.randperm(allrays.shape[0], generator=g)[:TRAIN_RAYS]].contiguous().to(dev)
    target = torch.rand(TRAIN_RAYS, 3, generator=g).to(dev)
    mask_gt = (torch.rand(TRAIN_RAYS, generator=g) > 0.7).float().to(dev)
    losses = []

    def one():
        ddp.zero_grad()
        r = render_rays(models, emb, rays, N_SAMPLES, False, 1.0, 1.0, N_IMPORTANCE, 32768, False, test_time=False,
                        compute_normal=True)
        r["_rays_d"] = rays[:, 3:6]
        loss = train_loss(r, target, mask_gt)
        loss.backward()
        ddp.step()
        losses.append(loss.detach())

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    points = TRAIN_RAYS * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE)
    return {"metric": "rays/sec (hash-grid train step: forward + backward + Adam; 4096-ray batch, 64+128 samples, analytic "
                      "normals, fp32)",
            "value": TRAIN_RAYS / ms * 1e3, "unit": "rays/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "dtype": "f32 (CUDA cores; gather / scatter-atomic bound)", "launches_per_step": (_lib.launch_count() - l0) / steps,
            "points_per_step": points, "table_atomics_per_step": points * 16 * 8 * 2,
            "parameters": int(ddp.flat_params.numel()), "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
            "note": "parity unpinned for the encoder (DESIGN.md 3.5); gradients pinned to the oracle's autograd "
                    "(tests/test_gpu_hashgrid.py, tests/test_hash_train_emu.py)"}


def hash_level_bench(dev, steps):
    """BASELINE config 3 shape: one eval render level (64+128 samples) of an 800x800 view with the hash-grid field
    (nerf_tcnn family; synthetic table and weights).  Returns a dict (rays/s per level)."""
    import torch
    from mirror_nerf_b200.mirror_nerf import Embedding
    from mirror_nerf_b200.rendering import render_rays
    from mirror_nerf_b200.synthetic import camera_rays
    models = hash_models(dev)
    emb = {"xyz": Embedding(0), "dir": Embedding(0)}
    c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.9]])
    rays = camera_rays(H, W, c2w=c2w, near=0.05, far=2.0).to(dev)
    fn = lambda: render_rays(models, emb, rays, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False, test_time=True,
                             compute_normal=False)
    with torch.no_grad():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n = rays.shape[0]
    return {"metric": "rays/sec per render level (hash-grid field, 64c+128f samples, eval, predicted normals)",
            "value": n / ms * 1e3, "unit": "rays/s", "ms_per_level": ms, "dtype": "f32 (CUDA cores; gather-bound)",
            "points_per_s": n * 256 / ms * 1e3, "table_read_GBps_algorithmic": n * 256 * 1024 / ms / 1e6,
            "note": "parity unpinned: tinycudann is an un-vendored dependency of the reference (DESIGN.md 3.5)"}


def cpu_train_rate(n_rays, threads):
    """The reference's training step (oracle port: torch CPU autograd, same loss) on the host cores."""
    import torch
    from mirror_nerf_b200.synthetic import camera_rays, scene_state_dicts
    from oracle import mirror_nerf_oracle as O
    torch.set_num_threads(threads)
    params = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in scene_state_dicts().items()}
    g = torch.Generator().manual_seed(1234)
    allrays = camera_rays(H, W, c2w=view_pose(0))
    rays = allrays[torch.randperm(allrays.shape[0], generator=g)[:n_rays]].contiguous()
    target = torch.rand(n_rays, 3, generator=g)
    mask_gt = (torch.rand(n_rays, generator=g) > 0.7).float()
    t0 = time.perf_counter()
    r = O.render_rays(params, rays, N_SAMPLES, False, 1.0, 1.0, N_IMPORTANCE, 32768, False, test_time=False,
                      compute_normal=True)
    r["_rays_d"] = rays[:, 3:6]
    train_loss(r, target, mask_gt).backward()
    return n_rays / (time.perf_counter() - t0)


def torch_cuda_baseline(dev):
    """Informational: the reference's PyTorch arithmetic (oracle restatement, fp32, TF32 off) on the same B200 --
    what the reference itself would do on this GPU.  Eval: 16,384 rays of the bench view, 1 bounce; train: one 1024-ray step."""
    import torch
    from mirror_nerf_b200.synthetic import camera_rays, scene_state_dicts
    from oracle import mirror_nerf_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    params = scene_state_dicts(device=dev)
    allrays = camera_rays(H, W, c2w=view_pose(0))
    rays = allrays[torch.linspace(0, allrays.shape[0] - 1, 16384).long()].contiguous().to(dev)
    fn = lambda r: O.render_rays(params, r, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False, test_time=True,
                                 compute_normal=False)
    with torch.no_grad():
        O.trace_eval(fn, rays[:2048], 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.trace_eval(fn, rays, 1)
        torch.cuda.synchronize()
        eval_rate = rays.shape[0] / (time.perf_counter() - t0)
    tp = {t: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for t, sd in params.items()}
    g = torch.Generator().manual_seed(1234)
    n = 1024
    tr = allrays[torch.randperm(allrays.shape[0], generator=g)[:n]].contiguous().to(dev)
    target, mask_gt = torch.rand(n, 3, generator=g).to(dev), (torch.rand(n, generator=g) > 0.7).float().to(dev)
    rates = []
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = O.render_rays(tp, tr, N_SAMPLES, False, 1.0, 1.0, N_IMPORTANCE, 32768, False, test_time=False, compute_normal=True)
        r["_rays_d"] = tr[:, 3:6]
        train_loss(r, target, mask_gt).backward()
        torch.cuda.synchronize()
        rates.append(n / (time.perf_counter() - t0))
    return {"eval_rays_per_s": eval_rate, "train_rays_per_s": rates[-1], "dtype": "f32 (torch CUDA, allow_tf32=False)",
            "note": "oracle restatement of the reference's torch ops on this GPU; informational, not the reference arm"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mirror_nerf_b200 import _lib
    from mirror_nerf_b200.mirror_nerf import Embedding, MirrorNeRF
    from mirror_nerf_b200.synthetic import camera_rays, scene_state_dicts
    from mirror_nerf_b200.trace import render_rays_recursive
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    models = {}
    for k, sd in scene_state_dicts().items():
        m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
        m.load_state_dict(sd)
        models[k] = m.to(dev).eval()
    emb = {"xyz": Embedding(10), "dir": Embedding(4)}
    rays_host = camera_rays(H, W, c2w=view_pose(rank)).pin_memory()
    rays_dev = rays_host.to(dev)
    n = rays_dev.shape[0]
    kw = dict(field_impl=args.field_impl)

    def step(rays):
        # the device-side recursion (mnrf_render_recursive): one C call per image, no host sync between level 0 and the blend
        return render_rays_recursive(models, emb, rays, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                     max_recursive_level=1, compact_outputs=not args.python_recursion, **kw,
                                     **({} if args.python_recursion else {"early_termination_eps": args.early_termination_eps}))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.fill_(1)  # L2 flush between timed iterations (outside the timed events)
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for _ in range(args.warmup):
            res = step(rays_dev)
        torch.cuda.synchronize()
        mirror_frac = float((res["mirror_mask_fine"] != 0).float().mean())

        # ---- device-resident throughput (value) + live kernel timing (roofline) ----
        sampler = ClockSampler(local)
        sampler.start()
        lib.mnrf_profile_enable(1)
        l0 = _lib.launch_count()
        ms_total = timed(lambda: step(rays_dev), args.steps)
        launches = _lib.launch_count() - l0
        lib.mnrf_profile_enable(0)
        k_ms, k_fl, k_n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(lib.mnrf_profile_collect(C.byref(k_ms), C.byref(k_fl), C.byref(k_n)))
        clocks = sampler.stop()

        # ---- end to end through the public API with host buffers: H2D of the rays from pinned memory, render, D2H of EVERY
        # tensor of the result dict into pinned memory (the reference moves every result tensor to the host, R/eval.py:735-736)
        res0 = step(rays_dev)
        out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in res0.items()}
        d2h_bytes = sum(v.numel() * v.element_size() for v in res0.values())

        def e2e_step():
            r = rays_host.to(dev, non_blocking=True)
            out = step(r)
            for k, v in out.items():
                out_host[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_steps = args.steps
        e2e_step()
        ms_e2e = timed(e2e_step, e2e_steps)

        # the same with the reference's FULL per-level dict (per-sample weights, z_vals, pred_normal: 4.4 KB per ray) through
        # the drop-in render_rays + the per-level Python driver; secondary, bounded to 2 steps (2.8 GB of D2H per step)
        e2e_full = None
        if world == 1 and not args.no_full_dict:
            def full_step(r):
                return render_rays_recursive(models, emb, r, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                             max_recursive_level=1, consume_rng=False, **kw)
            resf = full_step(rays_dev)
            full_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in resf.items()}
            full_bytes = sum(v.numel() * v.element_size() for v in resf.values())
            del resf

            def e2e_full_step():
                out = full_step(rays_host.to(dev, non_blocking=True))
                for k, v in out.items():
                    full_host[k].copy_(v, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            e2e_full_step()
            ms_full = timed(e2e_full_step, 2)
            e2e_full = {"value": n * 2 / (ms_full * 1e-3), "unit": "rays/s", "steps": 2, "h2d_bytes_per_step": n * 32,
                        "d2h_bytes_per_step": full_bytes, "keys": len(full_host),
                        "what": "render_rays (drop-in, full output dict) + per-level driver, every tensor copied to pinned host memory"}
            del full_host
            torch.cuda.empty_cache()

    other = None
    if args.field_impl != "tc3" and world == 1:
        # the library's default parity mode (three fp16 passes) on the same workload, for context
        def step3():
            return render_rays_recursive(models, emb, rays_dev, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                         max_recursive_level=1, compact_outputs=not args.python_recursion, field_impl="tc3",
                                         **({} if args.python_recursion else {"early_termination_eps": args.early_termination_eps}))
        with torch.no_grad():
            step3()
            ms3 = timed(step3, max(2, min(args.steps, 5)))
        other = {"field_impl": "tc3", "value": n * max(2, min(args.steps, 5)) / (ms3 * 1e-3), "unit": "rays/s",
                 "what": "same workload with the library default (3 fp16 passes, operand error 2^-21)"}
    value = world * n * args.steps / (ms_total * 1e-3)
    e2e_value = world * n * e2e_steps / (ms_e2e * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    achieved = (k_fl.value / 1e12) / (k_ms.value * 1e-3) if k_ms.value > 0 else None
    # DRAM bytes per launch of the dominant kernel: from the committed ncu capture of this same command (not measurable live)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_field_tc_traffic.json")))
        if tj.get("field_impl") == args.field_impl:   # a capture of another kernel variant says nothing about this run
            traffic, traffic_src = tj["dram_bytes_per_launch_avg"], tj["source"]
    except Exception:
        pass
    mma_per_mac = {"tc3": 3, "tc2": 2, "tc1": 1}[args.field_impl]

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"tc3": "f16x3-split operands, f32 accumulate (fp32-grade)",
                  "tc2": "f16 + 2 x e4m3 correction passes (fp8 datapath), f32 accumulate (operand error ~2^-16)",
                  "tc1": "f16, f32 accumulate"}[args.field_impl],
        "data": "synthetic",
        "config": CONFIG,
        "run": {"rays_per_step_per_gpu": n, "field_impl": args.field_impl, "mirror_ray_fraction": mirror_frac,
                "recursion": "per-level Python driver" if args.python_recursion else "mnrf_render_recursive (device-side, no host sync)",
                "early_termination_eps": 0.0 if args.python_recursion else args.early_termination_eps,
                "l2": "256 MB buffer written between timed steps (L2 flush); per-step scratch is > L2 anyway",
                "parallelism": f"ray-parallel x{world}, one view per rank, no collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps, "d2h": "all %d tensors of the compact per-ray result" % len(out_host)},
        "e2e_full_dict": e2e_full,
        "parity_mode": other,
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": (achieved / peak if achieved else None), "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "k_field_tc", "kernel_launches": int(k_n.value), "kernel_ms": k_ms.value,
                     "kernel_share_of_step": k_ms.value / ms_total if ms_total else None,
                     "flops_basis": "algorithmic 2*MAC of the reference layers (SURVEY.md 8d); the kernel issues "
                                    f"{mma_per_mac} fp16-pass equivalents of tensor-core work per algorithmic MAC",
                     "tensor_pipe_flops_frac": (achieved * mma_per_mac / peak if achieved else None),
                     "executed_frac": 0.95,
                     "executed_frac_note": "the kernel executes ~95 % of the algorithmic MACs: the sigma-only coarse pass skips "
                                           "normal_net (the reference computes and discards it) and the activation-free "
                                           "normal_net is folded to one 256->3 map; early termination (run.early_termination_eps) "
                                           "skips further fine chunks on scenes with opaque surfaces (config.field is translucent)"},
    }

    if not args.no_config4:
        # BASELINE config 4: 2 reflection bounces + roughness cone (8 jittered reflected rays = trace_ray_times 7), ray-parallel:
        # one 800x800 view per rank through mnrf_render_recursive (the 8 reflections of a level are ONE child batch)
        def step4():
            return render_rays_recursive(models, emb, rays_dev, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                         max_recursive_level=2, normal_noise_std=0.05, trace_ray_times=7, compact_outputs=True,
                                         early_termination_eps=args.early_termination_eps, with_level_rays=True, **kw)
        try:
            with torch.no_grad():
                l0 = _lib.launch_count()
                r4 = step4()
                per_step = _lib.launch_count() - l0
                level_rays = [int(v) for v in r4["level_rays"].cpu().tolist()]
                del r4
                ms4 = timed(step4, 2)
            line["config4"] = {"metric": "rays/sec (64c+128f samples, 2 bounces + roughness cone of 8 jittered reflections)",
                               "value": world * n * 2 / (ms4 * 1e-3), "unit": "rays/s", "ms_per_step": ms4 / 2, "steps": 2,
                               "n_gpus": world, "gpu_launches_per_step": per_step, "normal_noise_std": 0.05, "trace_ray_times": 7,
                               "rays_rendered_per_level": level_rays,
                               "ray_levels_per_s": world * sum(level_rays) * 2 / (ms4 * 1e-3),
                               "recursion": "mnrf_render_recursive (device-side)"}
        except Exception as e:
            line["config4"] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()

    if world > 1:
        # strong scaling: ONE 800x800 frame (rank 0's view) cut into tile-aligned shards (parallel.shard_bounds), every rank renders
        # its shard, rgb + depth are all-gathered (the only exchange); time = barrier-to-barrier max over ranks incl. the gather
        from mirror_nerf_b200.parallel import render_sharded
        frame = camera_rays(H, W, c2w=view_pose(0)).to(dev)

        def strong_step():
            render_sharded(lambda r: step(r.contiguous()), frame, rank, world, gather=("rgb_fine", "depth_fine"))
        with torch.no_grad():
            strong_step()
            ms_s = timed(strong_step, max(2, min(args.steps, 5)))
        ks = max(2, min(args.steps, 5))
        line["strong"] = {"metric": METRIC + ", ONE frame sharded over the ranks", "value": n * ks / (ms_s * 1e-3), "unit": "rays/s",
                          "n_gpus": world, "steps": ks, "ms_per_frame": ms_s / ks, "scaling": "strong",
                          "rays_per_rank": n // world, "gather": "all_gather of rgb_fine + depth_fine (16 B per ray, NCCL)"}

    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    if not args.no_train:
        ts = train_bench(dev, world, rank, max(2, min(args.steps, 5)), 3, args.peer_fused)
        # roofline of the training step: tensor bound on the algorithmic 1.77 GFLOP per ray; the layer GEMMs run
        # three tf32 passes (= 6 fp16-pass equivalents per MAC) and stream every activation through HBM once per layer
        tf = ts["algorithmic_tflops"]
        ts["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                          "traffic": None, "kernel": "k_gemm_tc_nn / k_gemm_tc_tn (unfused layer GEMMs, kind::tf32 x3)",
                          "flops_basis": "algorithmic 1.77 GFLOP per ray x 4096 rays per step / step time (whole step, not one kernel)",
                          "peak_source": peak_src,
                          "hbm_note": "ncu (profiles/r01_v3_train_gemm_tc_nn_ncu_metrics.txt): 1.0 GB read + 0.78 GB written per "
                                      "538 us layer launch (cold, under the profiler) = 3.3 TB/s = 51 % of the measured HBM peak; "
                                      "tensor pipe 81 % active; see roofline_gemm for the dominant kernel timed alone"}
        if rank == 0:
            try:
                # timed alone in short bursts: the burst figure of MEASURED_PEAKS.json is the denominator here
                burst = peaks.get("bf16_tflops") or peak
                gh = train_gemm_rooflines(dev, hbm_peak, burst)
                nn = gh["k_gemm_tc_nn"]
                # the dominant kernel of the step (k_gemm_tc_nn: 80 of 263 launches, 60 % of the kernel time), timed alone, against
                # BOTH of its bounds; the larger fraction is the one that binds
                hb = nn["hbm_frac"] >= nn["tensor_frac"]
                ts["roofline_gemm"] = {"bound": "hbm" if hb else "tensor",
                                       "achieved": nn["hbm_GBps"] if hb else nn["tf32_tflops_issued"],
                                       "peak": hbm_peak if hb else burst / 2.0, "unit": "GB/s" if hb else "TFLOP/s",
                                       "frac": max(nn["hbm_frac"], nn["tensor_frac"]), "traffic": None,
                                       "kernel": "k_gemm_tc_nn<1> (256 x 256 layer, bias + ReLU epilogue, 786,432 points)",
                                       "basis": "hbm: the fp32 activation matrix read once + the fp32 output written once (2 x 805 MB, "
                                                "algorithmic) / launch time; tensor: 3 tf32 passes x 2 x 786,432 x 256 x 256 flop / "
                                                "launch time against half the measured burst bf16 rate (MEASURED_PEAKS.json bf16_tflops; "
                                                "kind::tf32 runs at half the 16-bit rate); launch timed alone with CUDA events",
                                       "per_kernel": gh,
                                       "note": "knock-out timings of the same launch (tools/gemm_bench.py, profiles/r02_v5_train_gemm_"
                                               "operand_split_experiment.txt): 0.376 ms as is, 0.328 ms without the MMAs, 0.327 ms without "
                                               "the epilogue's global traffic, 0.296 ms with the MMAs alone -- the unfused layer GEMM sits at "
                                               "its HBM / tensor balance point with both overlapped; the lever is fusing consecutive layers "
                                               "per tile so that activations stop streaming through HBM (DESIGN.md 7)"}
            except Exception as e:
                ts["roofline_gemm"] = {"unavailable": repr(e)[:300]}
        line["train_step"] = ts
        if rank == 0:
            hl = hash_level_bench(dev, 3)
            a = hl["table_read_GBps_algorithmic"]
            hl["roofline"] = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak, "traffic": None,
                              "kernel": "k_field_hash", "bytes_basis": "algorithmic 1,024 B of table reads per point "
                              "(16 levels x 8 corners x 2 features x 4 B) x points per launch / launch time",
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6.65 TB/s",
                              "note": "the 46.5 MB table is L2-resident (ncu: 96.6 % L2 hit, 190 MB DRAM reads per 123 M-point launch); "
                                      "what bounds the kernel is the L1/TEX request rate of the gathers (ncu l1tex throughput 77 %, "
                                      "profiles/r01_v3_field_hash_ncu_metrics.txt), so the HBM fraction is a lower bound on how close "
                                      "the kernel is to ITS limit"}
            line["hash_grid_level"] = hl

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, rgb_ref, rays_s, _ = cpu_reference_rate(args.ref_rays, 3, 1, threads)
        with torch.no_grad():
            got = step(rays_s.to(dev))["rgb_fine"].cpu()
        mse = float(((got - rgb_ref) ** 2).mean())
        import math
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"{args.ref_rays} rays of the same view, 64+128 samples, 1 bounce, 3 steps after a 256-ray "
                                          "warm-up (BASELINE.md section 4)",
                                "best_step_rays_per_s": cpu_reference_rate.best}
        line["psnr_vs_reference_db"] = (-10 * math.log10(mse) if mse > 0 else float("inf"))
        # PSNR on a scene-like field: the analytic room scene (room_scene.py), field = tests/golden/room_field.npz (this repo's
        # training path, tools/train_room.py); ours (1 bounce, eval semantics) and the oracle against the analytic ground truth
        try:
            import numpy as np
            from mirror_nerf_b200.room_scene import room_pose, trace_room
            from oracle import mirror_nerf_oracle as O
            z = np.load(os.path.join(ROOT, "tests", "golden", "room_field.npz"))
            sds = {"coarse": {}, "fine": {}}
            for k in z.files:
                tag, name = k.split("/", 1)
                sds[tag][name] = torch.from_numpy(z[k].astype(np.float32))
            rmodels = {}
            for k, sd in sds.items():
                m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
                m.load_state_dict(sd)
                rmodels[k] = m.to(dev).eval()
            allr = camera_rays(400, 400, c2w=room_pose(2), near=0.05, far=12.0)
            rr = allr[torch.linspace(0, allr.shape[0] - 1, args.ref_rays).long()].contiguous()
            gt, gt_mask, _ = trace_room(rr)
            fn = lambda r: O.render_rays(sds, r, N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False, test_time=True,
                                         compute_normal=False)
            with torch.no_grad():
                ref_rgb = O.trace_eval(fn, rr, 1)["rgb_fine"]
                our_rgb = render_rays_recursive(rmodels, emb, rr.to(dev), N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                                max_recursive_level=1, compact_outputs=not args.python_recursion, **kw,
                                                **({} if args.python_recursion else {"early_termination_eps": args.early_termination_eps})
                                                )["rgb_fine"].cpu()
            ps = lambda x: -10 * math.log10(float(((x - gt) ** 2).mean()))

            def dist(a, b):   # relative error |a-b| / max(|b|, rms(b)): median / p99 / max / fraction beyond 1e-3
                a, b = a.double().flatten(), b.double().flatten()
                e = (a - b).abs() / b.abs().clamp_min(float(b.pow(2).mean().sqrt()))
                q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.double))
                return {"median": float(q[0]), "p99": float(q[1]), "max": float(e.max()), "frac_gt_1e-3": float((e > 1e-3).double().mean())}
            et = None
            if not args.python_recursion:   # what early termination skips on a scene with opaque surfaces (the room)
                with torch.no_grad():
                    st = render_rays_recursive(rmodels, emb, rr.to(dev), N_SAMPLES, False, 0, 0, N_IMPORTANCE, 32768, False,
                                               max_recursive_level=1, compact_outputs=True, with_stats=True,
                                               early_termination_eps=args.early_termination_eps, **kw)["fused_stats"].cpu().tolist()
                chunks = 2 * args.ref_rays * ((N_SAMPLES + N_IMPORTANCE + 31) // 32)
                et = {"eps": args.early_termination_eps, "chunks_skipped": int(st[1]), "chunks_total": chunks,
                      "skipped_frac": st[1] / chunks, "tiles_executed": int(st[0])}
            line["parity"] = {"rgb_fine": dist(our_rgb, ref_rgb), "rays": args.ref_rays, "field_impl": args.field_impl,
                              "early_termination_on_room_scene": et,
                              "against": "oracle port of the reference on the same rays (fitted room field, 1 bounce, eval semantics)",
                              "bar": "north-star: rgb within 1e-3 relative, PSNR within 0.05 dB"}
            line["psnr"] = {"ours_db": ps(our_rgb), "reference_db": ps(ref_rgb), "delta_db": ps(our_rgb) - ps(ref_rgb),
                            "ours_vs_reference_db": -10 * math.log10(max(float(((our_rgb - ref_rgb) ** 2).mean()), 1e-20)),
                            "rays": args.ref_rays, "mirror_ray_fraction": float(gt_mask.mean()),
                            "scene": "analytic box room with a planar mirror (mirror_nerf_b200/room_scene.py); field fitted by "
                                     "tools/train_room.py (6000 steps of this repo's training path), ground truth ray-traced"}
        except Exception as e:  # the fixture is optional
            line["psnr"] = {"unavailable": repr(e)}
        try:
            line["torch_cuda_baseline"] = torch_cuda_baseline(dev)
        except Exception as e:
            line["torch_cuda_baseline"] = {"unavailable": repr(e)[:200]}
        if "train_step" in line:
            line["train_step"]["cpu_baseline"] = {
                "value": cpu_train_rate(128, threads), "unit": "rays/s", "cores": threads, "kind": "port",
                "sample": "one 128-ray train step (forward + backward) of the oracle port on the host cores"}
    if rank == 0 and world == 1 and not args.no_train:
        try:  # last GPU work of the run: a failure here cannot touch the numbers above
            ht = hash_train_bench(dev, 3)
            # algorithmic bytes of a step: forward gathers (1,024 B per point) + backward recompute gathers (1,024 B) + table
            # gradient scatter (256 atomics x 4 B per point) per point
            by = ht["points_per_step"] * (1024 + 1024 + 1024) / (ht["ms_per_step"] * 1e-3) / 1e9
            ht["roofline"] = {"bound": "hbm", "achieved": by, "peak": hbm_peak, "unit": "GB/s", "frac": by / hbm_peak, "traffic": None,
                              "kernel": "k_field_hash + k_hash_bwd2",
                              "bytes_basis": "algorithmic 3,072 B per point (forward gathers, recompute gathers, gradient scatter) x "
                                             "1,048,576 points per step / step time",
                              "note": "issue / shared-memory-latency bound, not bandwidth bound: ncu of k_hash_bwd2 "
                                      "(profiles/r02_v1_hash_bwd2_ncu_metrics.txt): issue active 39.6 %, shared-memory wavefronts "
                                      "58 % of peak, short-scoreboard stalls 1.4 per issue, 12 % of the warp slots, table in L2"}
            line["hash_grid_train_step"] = ht
        except Exception as e:
            line["hash_grid_train_step"] = {"unavailable": repr(e)[:300]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to stdout with
    printf whenever NCCL_DEBUG >= VERSION is set in the environment): keep a private duplicate of fd 1 for the JSON line and point
    fd 1 at stderr for everything else in this process."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field-impl", default=os.environ.get("MNRF_BENCH_FIELD_IMPL", "tc2"), choices=["tc3", "tc2", "tc1"])
    ap.add_argument("--early-termination-eps", type=float, default=1e-5,
                    help="transmittance below which the fused fine pass stops a ray (0 = composite every sample)")
    ap.add_argument("--python-recursion", action="store_true",
                    help="drive the bounce from Python (per-level launches + host syncs) instead of mnrf_render_recursive")
    ap.add_argument("--ref-rays", type=int, default=4096, help="rays per step of the CPU reference sample (BASELINE.md section 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip BASELINE config 4 (2 bounces + roughness cone of 8 reflections)")
    ap.add_argument("--no-full-dict", action="store_true", help="skip the secondary end-to-end run with the reference's full output dict")
    ap.add_argument("--peer-fused", action="store_true",
                    help="train step: one peer-memory kernel (reduce-scatter + Adam + all-gather) instead of ncclAllReduce + Adam")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary train-step measurement (config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
