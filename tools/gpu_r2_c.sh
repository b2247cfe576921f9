#!/bin/bash
# round 2, call C (16 epilogue warps, one K chunk per column group): full GPU test suite on the new tree, tc2 vs tc3 bench, ncu of the tc2 kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_r2c.log
timeout 200 python bench.py --field-impl tc3 --no-train --no-cpu-baseline --steps 5 > gpurun_out/bench_r2c_tc3.json 2> gpurun_out/bench_r2c_tc3.err; echo "bench tc3 rc=$?"
timeout 200 python bench.py --field-impl tc2 --no-train --no-cpu-baseline --steps 5 > gpurun_out/bench_r2c_tc2.json 2> gpurun_out/bench_r2c_tc2.err; echo "bench tc2 rc=$?"
timeout 200 python bench.py --field-impl tc3 --no-train --no-cpu-baseline --steps 5 --early-termination-eps 0 > gpurun_out/bench_r2c_tc3_noet.json 2> gpurun_out/bench_r2c_tc3_noet.err; echo "bench tc3 no-ET rc=$?"
timeout 200 python bench.py --field-impl tc3 --no-train --no-cpu-baseline --steps 5 --python-recursion > gpurun_out/bench_r2c_tc3_py.json 2> gpurun_out/bench_r2c_tc3_py.err; echo "bench tc3 py rc=$?"
python - <<'PY'
import json
for f in ("tc3","tc2","tc3_noet","tc3_py"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r2c_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 ncu --set full -k regex:k_field_tc -s 4 -c 2 --clock-control none --import-source on -f -o gpurun_out/prof_r2c_tc2 python bench.py --field-impl tc2 --no-train --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2> gpurun_out/ncu_r2c_tc2.log; tail -2 gpurun_out/ncu_r2c_tc2.log
