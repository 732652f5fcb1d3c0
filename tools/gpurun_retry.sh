#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3; nothing is charged for those).  Usage:
#   tools/gpurun_retry.sh <log file> [gpurun args ...] -- '<command>'
LOG=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
