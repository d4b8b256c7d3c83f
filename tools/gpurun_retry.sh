#!/usr/bin/env bash
# gpurun with retries while the pod is busy (exit 3 / transient): usage gpurun_retry.sh <log> <timeout> <cmd...>
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|no box or slot" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
