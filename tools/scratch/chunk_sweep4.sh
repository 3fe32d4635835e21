#!/bin/bash
run() { python bench.py --no-extra --steps 5 --warmup 2 --realizations $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R', $1, 'threads', '$2', 'chunk steps', '$3', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
for R in 170 200 330; do
for cs in 0 20; do SMM_ABM_CHUNK_STEPS=$cs run $R auto $cs; done
done
SMM_ABM_THREADS=512 SMM_ABM_CHUNK_STEPS=20 run 170 512 20
SMM_ABM_THREADS=512 SMM_ABM_CHUNK_STEPS=20 run 200 512 20
