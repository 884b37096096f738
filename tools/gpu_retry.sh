#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   -- retries while the pool answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pool busy"; exit 3
