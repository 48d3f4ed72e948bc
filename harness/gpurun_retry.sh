#!/bin/bash
# usage: harness/gpurun_retry.sh <logfile> <timeout-seconds> <command string>
# retries while the pod answers "busy" (exit code 3); everything else is returned as is
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "rc=3 (gave up)" >> "$log"; exit 3
