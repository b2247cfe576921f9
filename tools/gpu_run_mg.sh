#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_mg$N.json 2> gpurun_out/bench_mg$N.err; echo "rc=$?" >> gpurun_out/bench_mg$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_mg${N}_ref.json 2>> gpurun_out/bench_mg$N.err
cat gpurun_out/bench_mg$N.json | cut -c 1-400; tail -3 gpurun_out/bench_mg$N.err; cut -c 1-200 gpurun_out/bench_mg${N}_ref.json
