#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   -- retries while the pod answers busy (rc 3 / transient)
log=$1; to=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
