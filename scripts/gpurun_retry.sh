#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <logfile> [--gpus N] -- <command>
# retries while the pod answers "busy" (exit 3: nothing charged), up to ~40 minutes
T=$1; LOG=$2; shift 2
EXTRA=()
while [ "$1" != "--" ]; do EXTRA+=("$1"); shift; done
shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${EXTRA[@]}" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
