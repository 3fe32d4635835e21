#!/bin/bash
# gpu_retry_n.sh NGPUS TIMEOUT 'command'
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "$3" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 120; continue; fi
  echo "$out" | tail -5; exit 0
done
echo "gave up"; exit 3
