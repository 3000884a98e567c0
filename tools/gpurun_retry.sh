#!/bin/bash
# gpurun with retries while the pod has no slot (exit code 3: nothing charged): tools/gpurun_retry.sh <timeout> '<command>'
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
