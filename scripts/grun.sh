#!/bin/bash
# gpurun with retries on "transient" (no GPU slot free).  usage: scripts/grun.sh <timeout-seconds> '<command>' [logfile]
T=$1; CMD=$2; LOG=${3:-/tmp/grun.log}
for i in $(seq 1 12); do
  gpurun --timeout "$T" -- "$CMD" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 45; continue; fi
  break
done
