#!/bin/bash
for ch in 2 4 2 4; do
SMM_ENSEMBLE_CHUNKS=$ch python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM%9)) bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['workloads']['abm_C2_one_call']
print('chunks per shard $ch', 'weak e2e ms', round(d['e2e']['ms_per_step'],2), 'one-call f64', round(o['full_us_float64']['ms_per_step'],2), 'f32', round(o['float32_densities']['ms_per_step'],2), 'means', round(o['means_only']['ms_per_step'],2))"
done
