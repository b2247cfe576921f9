#!/bin/bash
# round 2, call D: tests after the accumulator-overlap fix, the new bench line (all objects), weight-traffic experiment
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2d.log
for impl in tc3 tc2; do
  timeout 200 python bench.py --field-impl $impl --no-train --no-cpu-baseline --no-config4 --no-full-dict --steps 5 > gpurun_out/bench_r2d_$impl.json 2> gpurun_out/bench_r2d_$impl.err; echo "bench $impl rc=$?"
  MNRF_TC_DEBUG=1 timeout 200 python bench.py --field-impl $impl --no-train --no-cpu-baseline --no-config4 --no-full-dict --steps 5 > gpurun_out/bench_r2d_${impl}_quarterweights.json 2> gpurun_out/bench_r2d_${impl}_quarterweights.err; echo "bench $impl quarter weights rc=$?"
done
python - <<'PY'
import json
for f in ("tc3","tc3_quarterweights","tc2","tc2_quarterweights"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r2d_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 python bench.py --steps 5 > gpurun_out/bench_r2d_full.json 2> gpurun_out/bench_r2d_full.err; echo "full bench rc=$?"; tail -c 1500 gpurun_out/bench_r2d_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2d_ref.json 2> gpurun_out/bench_r2d_ref.err; echo "ref bench rc=$?"
