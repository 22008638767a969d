#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
# retries while gpurun answers 3 (no box / slot free right now; nothing charged)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
