#!/bin/bash
# usage: tools/gpuretry.sh <timeout> <command...>   retries while the pod is busy
t=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|busy\|draining"; then sleep 60; continue; fi
  echo "$out" | tail -40; exit 0
done
echo "gave up"; exit 3
