#!/bin/bash
# developer helper: retry a gpurun call while the pod answers "busy / draining" (nothing is charged for those)
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -qE "status=transient|status=busy|rc=None"; then sleep 90; continue; fi
  break
done
