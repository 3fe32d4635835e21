#!/bin/bash
nproc
for hw in 1 0; do
  echo "== SMM_HOST_WIDEN=$hw"
  SMM_HOST_WIDEN=$hw timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$hw bench.py --gpus 4 --steps 10 --warmup 3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
