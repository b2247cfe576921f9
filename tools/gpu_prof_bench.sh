#!/bin/bash
# ncu launch list of the bench command (kernel shares) + DRAM traffic of the dominant kernel
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_v3_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_v3_bench_under_ncu.json 2> gpurun_out/r01_v3_bench_under_ncu.err
tail -2 gpurun_out/r01_v3_bench_under_ncu.err; wc -l gpurun_out/r01_v3_bench_launches.csv
