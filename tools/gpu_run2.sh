#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?" >> gpurun_out/bench_tc3.err
timeout 600 python bench.py --steps 3 --warmup 3 --field-impl tc1 --no-cpu-baseline > gpurun_out/bench_tc1.json 2> gpurun_out/bench_tc1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 2 -f -o gpurun_out/prof_field_tc3 python tools/prof_one.py tc3 > gpurun_out/ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/pytest_all.log; cat gpurun_out/bench_tc3.json; tail -2 gpurun_out/bench_tc3.err; cat gpurun_out/bench_tc1.json gpurun_out/bench_ref.json; tail -2 gpurun_out/smoke.log
