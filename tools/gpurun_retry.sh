#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers busy (exit 3), up to 12 times
out=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
