#!/bin/bash
# usage: gpurun_retry.sh <timeout> <gpus> <command...>   — retries while the pod answers "transient/busy"
T=$1; G=$2; shift 2
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 150; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after retries"
