#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...> : retries while the pod answers busy (rc 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 60; continue; fi
  break
done
tail -40 /tmp/gpurun_last.log
