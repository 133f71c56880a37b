#!/bin/bash
# gpurun with retry on "busy" (exit 3).  Usage: scripts/gpurun_retry.sh <log-tag> <timeout-s> [--gpus N] -- '<command>'
tag=$1; shift
tmo=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout $tmo "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  echo "attempt $i exit $rc" >> gpurun_out/${tag}_call.log
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
