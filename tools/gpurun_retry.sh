#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]   -- retries while gpurun answers "no box free" (exit 3)
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
