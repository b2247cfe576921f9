#!/bin/bash
mkdir -p gpurun_out
for impl in tc3 tc1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 2 -f -o gpurun_out/prof_v2_$impl python tools/prof_one.py $impl > gpurun_out/ncu_$impl.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
