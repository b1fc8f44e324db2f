#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout_s> <command...>  — retries while the pod answers busy (exit 3 / transient)
gpus=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  if [ "$gpus" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $to -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" 2>&1); fi
  rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy, sleeping 90 s"; sleep 90; continue; fi
  exit $rc
done
exit 3
