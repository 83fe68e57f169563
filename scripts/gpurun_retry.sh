#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  else out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1); rc=$?; fi
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes\|retry later"; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
