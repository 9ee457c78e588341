#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...]   - retries while the pod answers "busy" (exit code 3)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i attempt(s)"; exit $rc; fi
  sleep 90
done
echo "gave up (busy)"; exit 3
