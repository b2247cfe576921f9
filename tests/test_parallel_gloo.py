"""world_size-2 (and 3) gloo run of the ray-sharding host logic on CPU: every rank renders its shard with a stand-in
render function and the gathered image equals the single-process result, independent of the split."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(rays):
    # deterministic per-ray function (a ray's result must not depend on which shard it lands in)
    rgb = torch.sin(rays[:, :3] * 3.0 + rays[:, 3:6])
    return {"rgb_fine": rgb, "depth_fine": rays[:, 6] + rays[:, 7] * 0.5}


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mirror_nerf_b200.parallel import render_sharded, shard_bounds
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(n, 8, generator=g)
    local, full = render_sharded(_fake_render, rays, rank, world)
    lo, hi = shard_bounds(n, rank, world)
    assert local["rgb_fine"].shape[0] == hi - lo
    want = _fake_render(rays)
    assert torch.equal(full["rgb_fine"], want["rgb_fine"])
    assert torch.equal(full["depth_fine"], want["depth_fine"])
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def _run(world, n, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_sharded_render_world2(tmp_path):
    _run(2, 1000, tmp_path)


def test_sharded_render_world3_ragged(tmp_path):
    _run(3, 130, tmp_path)  # 2 tiles over 3 ranks: one rank gets an empty shard


# ---- data-parallel training plumbing (flat gradient all-reduce), world_size 2 over gloo ---------------------------------
def _ddp_worker(rank, world, port, out_dir, family="mlp"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mirror_nerf_b200.mirror_nerf import MirrorNeRF
    from mirror_nerf_b200.parallel import FlatDataParallel
    from mirror_nerf_b200.synthetic import make_state_dict
    models = {}
    for name, seed in (("coarse", 1), ("fine", 2)):
        if family == "hash":  # nerf_tcnn family: the hash tables join the same flat all-reduce
            from mirror_nerf_b200.mirror_nerf_tcnn import MirrorNeRFTcnn
            torch.manual_seed(seed)
            m = MirrorNeRFTcnn(bound=1, predict_normal=True, predict_mirror_mask=True)
        else:
            m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True)
            m.load_state_dict(make_state_dict(seed))
        models[name] = m
    before = {k: {n: p.detach().clone() for n, p in m.named_parameters()} for k, m in models.items()}
    ddp = FlatDataParallel(models)
    if family == "hash":
        assert ddp.flat_params.numel() == 2 * (12196240 + 11041) == ddp.flat_grads.numel()  # 97.7 MB per step
    else:
        assert ddp.flat_params.numel() == 2 * 662152 == ddp.flat_grads.numel()  # 5.30 MB per step
    for k, m in models.items():  # flattening keeps values and makes every parameter / grad a view of the flat buffers
        for n, p in m.named_parameters():
            assert torch.equal(p.detach(), before[k][n])
            assert p.grad is not None and p.grad.shape == p.shape
    lo, hi = ddp.flat_params.data_ptr(), ddp.flat_params.data_ptr() + 4 * ddp.flat_params.numel()
    assert all(lo <= p.data_ptr() < hi for p in ddp.params)
    # rank-dependent "gradients" written through autograd's accumulation path
    g = torch.Generator().manual_seed(100 + rank)
    loss = sum((p * torch.randn(p.shape, generator=g)).sum() for p in ddp.params)
    loss.backward()
    mine = ddp.flat_grads.clone()
    total = ddp.all_reduce_grads().clone()
    # reference: the same draws for both ranks, summed locally
    want = torch.zeros_like(total)
    for r in range(world):
        gg = torch.Generator().manual_seed(100 + r)
        want += torch.cat([torch.randn(p.shape, generator=gg).reshape(-1) for p in ddp.params])
    assert torch.allclose(total, want, rtol=0, atol=1e-6)
    assert not torch.equal(mine, total)
    ddp.zero_grad()
    assert float(ddp.flat_grads.abs().max()) == 0.0 and all(float(p.grad.abs().max()) == 0.0 for p in ddp.params)
    models["coarse"].zero_grad(set_to_none=True)
    ddp.zero_grad()
    assert all(p.grad is not None for p in ddp.params)
    try:
        ddp.step()
        raise AssertionError("step() must refuse to run without CUDA")
    except RuntimeError as e:
        assert "CUDA" in str(e)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ddp_ok{rank}_{family}"), "w").write("ok")


def test_flat_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ddp_ok{r}_mlp")) for r in range(2))


def test_flat_gradient_allreduce_world2_hash_grid_family(tmp_path):
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path), "hash"), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ddp_ok{r}_hash")) for r in range(2))
