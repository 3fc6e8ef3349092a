#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT 'command'   — retries while the pod answers busy (exit 3)
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
