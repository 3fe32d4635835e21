#!/bin/bash
# with step chunks the realizations need not all be resident: wider CTAs for mid-size ensembles?
run() { python bench.py --no-extra --steps 5 --warmup 2 --realizations $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R', $1, 'threads', '$2', 'chunk steps', '$3', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
for R in 370 512 600 700 1024; do
for th in 128 256; do SMM_ABM_THREADS=$th SMM_ABM_CHUNK_STEPS=20 run $R $th 20; done
done
SMM_ABM_THREADS=256 SMM_ABM_CHUNK_STEPS=0 run 370 256 0
