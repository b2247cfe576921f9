#!/bin/bash
# unrolled issue loop of the default tc2 schedule: regression, then the bench line with the unrolled (default) and the generic loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_issue_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_issue_pytest.log
for d in 0 2 0 2; do
  MNRF_TC_DEBUG=$d timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-config4 --no-full-dict --no-cpu-baseline \
    > gpurun_out/r02_issue_bench_d$d.json 2> gpurun_out/r02_issue_bench_d$d.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_issue_bench_d$d.json"))
    print("MNRF_TC_DEBUG=$d value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
done
