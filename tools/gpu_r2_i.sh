#!/bin/bash
mkdir -p gpurun_out
MNRF_TC_DEBUG=1 MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc2 > gpurun_out/trace_r2i_tc2_quarterweights.txt 2>&1; echo "rc=$?"; head -6 gpurun_out/trace_r2i_tc2_quarterweights.txt
MNRF_TC_DEBUG=1 MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc3 > gpurun_out/trace_r2i_tc3_quarterweights.txt 2>&1; echo "rc=$?"; head -6 gpurun_out/trace_r2i_tc3_quarterweights.txt
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc3 > gpurun_out/trace_r2i_tc3.txt 2>&1; echo "rc=$?"; head -6 gpurun_out/trace_r2i_tc3.txt
