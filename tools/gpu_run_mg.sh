#!/bin/bash
# multi-GPU bench as the driver launches it (one rank per GPU over NCCL): weak scaling headline, strong-scaling object, config 4,
# train step with the flat NCCL all-reduce.   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_run_mg.sh N'
N=${1:-2}
STEPS=${2:-5}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/r02_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_${N}gpu_reference.json 2>> gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","gpu_launches","scaling")}); print("e2e", d["e2e"]["value"]); print("strong", d.get("strong"))
print("config4", {k:d["config4"].get(k) for k in ("value","ms_per_step","gpu_launches_per_step","n_gpus")} if "config4" in d else None)
t=d.get("train_step") or {}; print("train", {k:t.get(k) for k in ("value","ms_per_step","exchange","allreduce_bytes_per_step")})
PY
tail -3 gpurun_out/r02_bench_${N}gpu.err; cut -c 1-160 gpurun_out/r02_bench_${N}gpu_reference.json
