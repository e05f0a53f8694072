#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   (retries while the pod answers "busy")
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.out 2>&1
  if grep -q "status=transient" /tmp/gpurun_last.out; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.out
