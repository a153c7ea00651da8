#!/bin/bash
# usage: gpuretry.sh <tries> <gpurun args...>   retries while gpurun answers 3 (busy, nothing charged)
tries=$1; shift
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpuretry] attempt $i busy, sleeping"; sleep 90
done
exit 3
