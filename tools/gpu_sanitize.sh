#!/bin/bash
# compute-sanitizer over small configurations of every path (memcheck + initcheck on the hash field)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_hashgrid.py -m gpu -q -x -k "without_heads or render_rays" > gpurun_out/sanitizer_hash_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_hash_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_ray or variants_golden and tc3 and white_back or composite or reflect_compact" > gpurun_out/sanitizer_eval_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_eval_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "odd_sizes or ray_gradients" > gpurun_out/sanitizer_train_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_train_memcheck.log
