#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'  [gpus]
# Retries while the pod answers "transient" (busy: nothing charged), up to 40 times.
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] busy"; sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
