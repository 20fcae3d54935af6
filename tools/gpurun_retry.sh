#!/bin/bash
# gpurun_retry.sh LOGFILE [gpurun args...] : retries a gpurun call while the pod answers "busy" (exit code 3, nothing charged)
log=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
