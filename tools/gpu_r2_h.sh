#!/bin/bash
# round 2, call H: spinning waits for producer and issuer: timeline, tests, bench
mkdir -p gpurun_out
MNRF_LIB=$PWD/mirror_nerf_b200/lib/libmnrf_trace.so timeout 120 python tools/tc_trace.py tc2 > gpurun_out/trace_r2h_tc2.txt 2>&1; echo "trace tc2 rc=$?"; head -6 gpurun_out/trace_r2h_tc2.txt
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2h.log
for impl in tc3 tc2; do
  timeout 200 python bench.py --field-impl $impl --no-train --no-cpu-baseline --no-config4 --no-full-dict --steps 5 > gpurun_out/bench_r2h_$impl.json 2> gpurun_out/bench_r2h_$impl.err; echo "bench $impl rc=$?"
done
python - <<'PY'
import json
for f in ("tc3","tc2"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r2h_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
