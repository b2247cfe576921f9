#!/bin/bash
mkdir -p gpurun_out
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc2 > gpurun_out/trace_r2f_tc2.txt 2>&1; echo "trace tc2 rc=$?"; head -6 gpurun_out/trace_r2f_tc2.txt
