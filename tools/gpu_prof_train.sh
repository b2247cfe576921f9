#!/bin/bash
# launch list of one train step + full ncu capture of the dominant training GEMM
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/train_launches.csv python tools/train_perf.py 1024 > gpurun_out/train_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_nn -s 10 -c 2 -f -o gpurun_out/prof_train_gemm_nn python tools/train_perf.py 4096 > gpurun_out/ncu_train_nn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn -s 4 -c 1 -f -o gpurun_out/prof_train_gemm_tn python tools/train_perf.py 4096 > gpurun_out/ncu_train_tn.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/train_launches.csv; tail -2 gpurun_out/train_launches.log
