#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3):
#   tools/gpurun_retry.sh <timeout-s> [--gpus N] <command>
t=$1; shift
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t "${extra[@]}" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
