#!/bin/bash
# usage: gpu_retry.sh <timeout> <cmd...>; retries on rc=3 (no box) up to 12 times
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
