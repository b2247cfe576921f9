#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?" >> gpurun_out/bench_tc3.err
grep -E "gradient parity|passed|failed|Error" gpurun_out/pytest_train.log | tail -30; python -c "
import json; d=json.load(open('gpurun_out/bench_tc3.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d.get('train_step'), indent=1))"; tail -3 gpurun_out/bench_tc3.err
