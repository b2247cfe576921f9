#!/bin/bash
# full ncu captures of the training GEMM kernels (one launch each of the fine pass)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_nn -s 25 -c 1 -f -o gpurun_out/prof_tc_nn python tools/train_perf.py 4096 > gpurun_out/ncu_tc_nn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_tn -s 30 -c 1 -f -o gpurun_out/prof_tc_tn python tools/train_perf.py 4096 > gpurun_out/ncu_tc_tn.log 2>&1
ls -la gpurun_out/prof_tc_*.ncu-rep; tail -2 gpurun_out/ncu_tc_nn.log
