#!/bin/bash
# block until no gpurun call from this repo is in flight (max ~25 min)
for i in $(seq 1 100); do
  if /usr/local/graft/bin/gpurun --status 2>/dev/null | grep -q '"in_flight": 0'; then exit 0; fi
  sleep 15
done
exit 1
