#!/bin/bash
# N-split schedule of the tc2 field kernels: regression under both schedules, then the bench line for
# MNRF_TC_SPLIT = 0 (unsplit), 1 (split, two column groups converting at a time), 2 (split, all four at once)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r02_split_pytest.log 2>&1
echo "pytest split rc=$?"; tail -3 gpurun_out/r02_split_pytest.log
for s in 0 1 2; do
  MNRF_TC_SPLIT=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-config4 --no-full-dict --no-cpu-baseline \
    > gpurun_out/r02_split_bench_s$s.json 2> gpurun_out/r02_split_bench_s$s.err
  echo "bench split=$s rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_split_bench_s$s.json"))
    print("split=$s value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "parity_mode tc3", round(d["parity_mode"]["value"]), d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
done
