#!/bin/bash
# round 2, final regression + evidence on one B200 (after the packed-fp32 epilogue / role placement / N-split option):
#   gpurun --timeout 3000 -- 'bash tools/gpu_r2_final2.sh'
mkdir -p gpurun_out
bash tools/gpu_r2_final.sh
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool:"; tail -2 gpurun_out/r02_sanitizer_$tool.log
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-config4 --no-full-dict > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
wc -l gpurun_out/r02_bench_launches.csv
