#!/bin/bash
# gpu_retry.sh TIMEOUT 'command' : retry gpurun while the pod answers busy (exit 3 / transient)
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $1 -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 90; continue; fi
  echo "$out" | tail -5; exit 0
done
echo "gave up"; exit 3
