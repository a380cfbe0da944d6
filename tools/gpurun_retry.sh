#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <logfile> [--gpus N] <command...>  -- retries while the pod answers "busy" (rc 3)
TO=$1; LOG=$2; shift 2
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$TO" $G -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc attempt=$attempt" >> "$LOG"; exit $rc; fi
  sleep 45
done
echo "gave up" >> "$LOG"; exit 3
