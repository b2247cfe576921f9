#!/bin/bash
# A/B of two library builds (mirror_nerf_b200/lib/libmnrf_old.so = the previous commit, libmnrf.so = the working tree): tc2 regression
# on the new one, then the bench line ABAB
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_ab_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_ab_pytest.log
for lib in libmnrf_old libmnrf libmnrf_old libmnrf; do
  MNRF_LIB=$PWD/mirror_nerf_b200/lib/$lib.so timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-config4 --no-full-dict --no-cpu-baseline \
    > gpurun_out/r02_ab_bench_$lib.json 2> gpurun_out/r02_ab_bench_$lib.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_ab_bench_$lib.json"))
    print("$lib value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
except Exception as e:
    print("no bench line", e)
PY
done
