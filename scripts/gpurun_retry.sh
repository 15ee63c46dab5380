#!/bin/bash
# usage: scripts/gpurun_retry.sh LOG TIMEOUT 'command' [--gpus N]
# retries while the pod answers "busy" (exit 3); everything else is final
LOG=$1; TO=$2; CMD=$3; shift 3
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" -- "$CMD" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $attempt attempt(s)" >> $LOG; exit $rc; fi
  sleep 90
done
