#!/bin/bash
# full GPU check: parity tests, smoke, default bench (tc3) + reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?" >> gpurun_out/bench_tc3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/pytest_all.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_tc3.json; tail -3 gpurun_out/bench_tc3.err; cat gpurun_out/bench_ref.json
