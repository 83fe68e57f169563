#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
