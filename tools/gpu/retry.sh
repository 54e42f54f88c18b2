#!/bin/bash
# usage: tools/gpu/retry.sh LOGFILE [gpurun args...] -- retries while the pod answers "busy" (exit 3), up to ~40 min
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 100
done
exit 3
