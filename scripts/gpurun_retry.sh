#!/bin/bash
# usage: scripts/gpurun_retry.sh <logname> <gpurun args...>   (retries while the pod answers "busy")
log=gpurun_out/$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "done rc=$rc" >> "$log"; exit $rc; fi
  sleep 45
done
echo "gave up" >> "$log"
