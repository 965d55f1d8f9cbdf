#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "transient"/busy
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy" "$log"; then sleep 150; else break; fi
done
