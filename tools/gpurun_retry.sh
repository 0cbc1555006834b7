#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / transient: nothing is charged for those).
#   tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
LOG=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
