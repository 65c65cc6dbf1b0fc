#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'   -- retries while the pod answers "busy" (exit 3)
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
