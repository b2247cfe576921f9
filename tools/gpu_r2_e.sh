#!/bin/bash
# round 2, call E: device timeline of the tc2 / tc3 kernels (trace flavour of the library), determinism of the normals kernel
mkdir -p gpurun_out
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc2 > gpurun_out/trace_r2e_tc2.txt 2>&1; echo "trace tc2 rc=$?"; head -8 gpurun_out/trace_r2e_tc2.txt
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc3 > gpurun_out/trace_r2e_tc3.txt 2>&1; echo "trace tc3 rc=$?"; head -8 gpurun_out/trace_r2e_tc3.txt
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc1 > gpurun_out/trace_r2e_tc1.txt 2>&1; echo "trace tc1 rc=$?"; head -8 gpurun_out/trace_r2e_tc1.txt
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "deterministic or train_mode or heads_optional or analytic_normal" > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r2e.log
