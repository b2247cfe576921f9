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
