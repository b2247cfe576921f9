#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 3 5 6 7; do echo "== MNRF_TC_DEBUG=$d"; MNRF_TC_DEBUG=$d timeout 200 python tools/gpu_diag.py perf 2>&1 | grep "rays/s"; done | tee gpurun_out/dbg.log
