#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [gpus]  — retries while the pod answers "busy" (exit 3) or while an earlier
# call of this repo is still registered (exit 2 with "already running"); nothing is charged for those
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 60); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > /tmp/gpurun_attempt.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > /tmp/gpurun_attempt.log 2>&1; fi
  rc=$?
  cat /tmp/gpurun_attempt.log
  if [ $rc -eq 3 ]; then echo "[retry] busy, attempt $i"; sleep 90; continue; fi
  if [ $rc -eq 2 ] && grep -q "already running" /tmp/gpurun_attempt.log; then echo "[retry] previous call still registered, attempt $i"; sleep 60; continue; fi
  exit $rc
done
exit 3
