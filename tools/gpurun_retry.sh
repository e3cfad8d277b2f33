#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / transient): tools/gpurun_retry.sh <log> <timeout_s> <command...>
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -5 $LOG | cut -c1-600
