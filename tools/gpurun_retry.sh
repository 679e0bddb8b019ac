#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> <command...>   — retries while the pod answers busy (exit code 3)
log="$1"; shift; to="$1"; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $attempt attempt(s)" >> "$log"; exit $rc; fi
  sleep 120
done
exit 3
