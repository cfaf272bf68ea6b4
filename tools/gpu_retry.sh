#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <command...>   retries while the pod answers busy (exit 3)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
