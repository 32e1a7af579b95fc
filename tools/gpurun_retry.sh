#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <log>: retries while the pod's GPU slots are busy (exit code 3 = nothing charged)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" > $3 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
