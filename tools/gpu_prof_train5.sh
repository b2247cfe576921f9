#!/bin/bash
# round-1 v3 evidence for the training path: sanitizer on a tiny step, launch list of one 4096-ray step, full captures of the two tensor-core GEMMs
mkdir -p gpurun_out
cat > /tmp/tiny_train.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from test_gpu_train import _models, _smooth_sds, _rng, _loss
from mirror_nerf_b200.rendering import render_rays
from mirror_nerf_b200.synthetic import random_rays
models, emb = _models(_smooth_sds())
rays = random_rays(6, seed=3).cuda().requires_grad_(True)
r = render_rays(models, emb, rays, 64, False, 1.0, 1.0, 128, 32768, False, test_time=False, compute_normal=True, rng=_rng(6))
_loss(r, rays[:, 3:6].detach(), 0).backward()
torch.cuda.synchronize()
print("tiny train step ok", float(rays.grad.abs().sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/tiny_train.py > gpurun_out/sanitizer_train_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_train_memcheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_v3_train_launches.csv python tools/train_perf.py 4096 > gpurun_out/r01_v3_train_launches.log 2>&1; tail -1 gpurun_out/r01_v3_train_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_nn -s 3 -c 1 -f -o gpurun_out/prof_tc_nn python tools/train_perf.py 4096 > gpurun_out/ncu_tc_nn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_tn -s 3 -c 1 -f -o gpurun_out/prof_tc_tn python tools/train_perf.py 4096 > gpurun_out/ncu_tc_tn.log 2>&1
ls -la gpurun_out/prof_tc_*.ncu-rep
