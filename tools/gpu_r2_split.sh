#!/bin/bash
# N-split schedule / role placement of the tc2 field kernels: full GPU regression, then the bench line for
# {default library, roles on the highest warp ids} x {unsplit, split}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_split_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_split_pytest.log
for lib in libmnrf libmnrf_hi; do
  [ -f mirror_nerf_b200/lib/$lib.so ] || continue
  for s in 0 1; do
    MNRF_LIB=$PWD/mirror_nerf_b200/lib/$lib.so MNRF_TC_SPLIT=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-train --no-config4 --no-full-dict --no-cpu-baseline \
      > gpurun_out/r02_split_bench_${lib}_s$s.json 2> gpurun_out/r02_split_bench_${lib}_s$s.err
    echo "bench $lib split=$s rc=$?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_split_bench_${lib}_s$s.json"))
    print("$lib split=$s value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "parity_mode tc3", round(d["parity_mode"]["value"]), d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
  done
done
