#!/bin/bash
# usage: scripts/gpurun_retry.sh <log file> <gpurun args...>   — retries while the pod answers "transient" (exit 3: nothing charged)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
echo "attempts=$attempt rc=$rc"
