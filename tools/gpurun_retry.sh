#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (no box / slot free: nothing is charged).
# Usage: tools/gpurun_retry.sh <log-file> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 150; continue; fi
  break
done
tail -80 "$LOG"
