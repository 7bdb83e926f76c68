#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- <command>     retries while the pod answers "busy" (exit code 3)
log=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 60
done
exit 3
