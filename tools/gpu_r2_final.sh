#!/bin/bash
# round 2 regression on one B200: GPU tests, smoke(), both bench arms as the driver runs them, initcheck, ncu of the fused kernel
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2_final.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","clocks")}); print("e2e", d["e2e"]); print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","traffic","kernel_share_of_step")})
print("parity_mode", d.get("parity_mode")); print("config4", d.get("config4")); print("parity", d.get("parity")); print("psnr", d.get("psnr"))
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_initcheck.log 2>&1; tail -2 gpurun_out/r02_sanitizer_initcheck.log
timeout 400 ncu --set full -k regex:k_field_tc -s 7 -c 1 --clock-control none --import-source on -f -o gpurun_out/r02_prof_field_tc2_fused python bench.py --field-impl tc2 --no-train --no-cpu-baseline --no-config4 --no-full-dict --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r02_ncu_tc2_fused.log; tail -1 gpurun_out/r02_ncu_tc2_fused.log
