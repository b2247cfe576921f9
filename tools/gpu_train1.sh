#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
timeout 300 python tools/train_perf.py 4096 > gpurun_out/train_perf.log 2>&1
tail -40 gpurun_out/pytest_train.log; cat gpurun_out/train_perf.log | tail -5
