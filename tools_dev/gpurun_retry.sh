#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3): tools_dev/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i"; sleep 60
done
exit 3
