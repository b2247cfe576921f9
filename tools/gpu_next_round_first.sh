#!/bin/bash
# First GPU call of the next round (plan item 1): re-measure the final tree of round 1.
#   gpurun --timeout 900 -- 'bash tools/gpu_next_round_first.sh'
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
# 8-warp hash-grid backward: memcheck (racecheck is in profiles/r01_v5_*), ncu full capture of the fine- and coarse-pass launches
timeout 120 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_hashgrid.py -m gpu -q -x -k "ray_gradients and True" > gpurun_out/sanitizer_hash_bwd2_memcheck.log 2>&1; tail -2 gpurun_out/sanitizer_hash_bwd2_memcheck.log
timeout 150 ncu --set full -k regex:k_hash_bwd2 -s 2 -c 2 --clock-control none --import-source on -f -o gpurun_out/prof_hash_bwd2 python tools/hash_train_perf.py > /dev/null 2> gpurun_out/ncu_hash_bwd2.log; tail -2 gpurun_out/ncu_hash_bwd2.log
MNRF_HASH_BWD_LAYOUT=32x1 timeout 60 python tools/hash_train_perf.py > gpurun_out/hash_train_perf_32x1.json 2>/dev/null
timeout 60 python tools/hash_train_perf.py > gpurun_out/hash_train_perf_16x2.json 2>/dev/null
grep -h ms_per_step gpurun_out/hash_train_perf_32x1.json gpurun_out/hash_train_perf_16x2.json
