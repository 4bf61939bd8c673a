#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3 / status=transient).  usage: tools/gpurun_retry.sh LOG TIMEOUT [--gpus N] -- 'cmd'
LOG=$1; shift
TMO=$1; shift
for k in $(seq 1 40); do
  gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
