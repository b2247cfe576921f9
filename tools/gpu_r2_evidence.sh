#!/bin/bash
# round 2 evidence: compute-sanitizer on the new kernels, ncu full captures of the fused field kernel (tc2, tc3), launch list of
# the bench command with DRAM bytes per launch.   gpurun --timeout 2400 -- 'bash tools/gpu_r2_evidence.sh'
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_initcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_initcheck.log
for impl in tc2 tc3; do
  timeout 400 ncu --set full -k regex:k_field_tc -s 6 -c 2 --clock-control none --import-source on -f -o gpurun_out/r02_prof_field_$impl python bench.py --field-impl $impl --no-train --no-cpu-baseline --no-config4 --no-full-dict --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r02_ncu_$impl.log; tail -1 gpurun_out/r02_ncu_$impl.log
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-config4 --no-full-dict > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
wc -l gpurun_out/r02_bench_launches.csv
