#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
timeout 300 python tools/gpu_diag.py perf > gpurun_out/diag_perf.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; echo "rc=$?" >> gpurun_out/bench_tc3.err
timeout 600 python bench.py --steps 3 --warmup 3 --field-impl tc1 --no-cpu-baseline > gpurun_out/bench_tc1.json 2> gpurun_out/bench_tc1.err
grep -E "passed|failed|^E  " gpurun_out/pytest_all.log | head -20; tail -3 gpurun_out/diag_perf.log; python - <<'PY'
import json
for f in ("gpurun_out/bench_tc3.json","gpurun_out/bench_tc1.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value %.0f e2e %.0f ms/step %.1f achieved %.1f TF frac %.3f tpipe %.3f psnr %s cpu %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["achieved"],r["frac"],r["tensor_pipe_flops_frac"],d.get("psnr_vs_reference_db"),d.get("cpu_baseline",{}).get("value")))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_tc3.err
