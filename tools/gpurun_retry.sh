#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout> <cmd...>
t=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout "$t" -- "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
