#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/pytest_train.log 2>&1; tail -3 gpurun_out/pytest_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/train_tc_launches.csv python tools/train_perf.py 4096 > gpurun_out/train_tc_launches.log 2>&1
tail -1 gpurun_out/train_tc_launches.log
