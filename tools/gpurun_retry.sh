#!/bin/bash
# gpurun with retries while the pod answers "no slot right now": gpurun_retry.sh [gpus=N] <timeout> <command...>
G=""
case "$1" in gpus=*) G="--gpus ${1#gpus=}"; shift;; esac
TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $G --timeout $TO -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
tail -25 /tmp/gpurun_last.log
