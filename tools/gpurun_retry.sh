#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3): tools/gpurun_retry.sh <timeout-s> <command...>
t=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  grep -q transient gpurun_out/.last_call.json 2>/dev/null
  sleep 120
done
exit 3
