#!/bin/bash
# gpurun with retries while the pod answers "no slot right now" (exit code 3): gpurun_retry.sh <timeout> <command...>
TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
tail -25 /tmp/gpurun_last.log
