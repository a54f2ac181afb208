#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: gpurun_retry.sh <timeout> [--gpus N] <cmd...>
t=$1; shift
opts=""
if [ "$1" = "--gpus" ]; then opts="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  gpurun --timeout "$t" $opts -- "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
