#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
#   profiles/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun args]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] busy, sleeping 120 s"; sleep 120
done
exit 3
