#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:k_gemm_tc -s 60 -c 60 --csv --log-file gpurun_out/train_tc_launches2.csv python tools/train_perf.py 4096 > gpurun_out/train_tc_launches2.log 2>&1
tail -1 gpurun_out/train_tc_launches2.log
