#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (exit code 3): bash tools/gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1; rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
