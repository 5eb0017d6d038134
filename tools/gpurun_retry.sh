#!/bin/bash
# tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command': retries while the pod answers "transient" (nothing charged)
LOG=$1; TMO=$2; shift 2
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 90
done
