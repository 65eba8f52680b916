#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers "transient" (nothing charged)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 150
done
