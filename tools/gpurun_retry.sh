#!/bin/bash
# usage: gpurun_retry.sh <timeout> [--gpus N] -- '<command>'   : retries while the pod has no free slot (exit code 3), up to ~60 min
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
