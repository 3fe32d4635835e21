#!/bin/bash
# step chunks at other ensemble sizes / CTA widths (automatic width): one chunk vs chunks of 20 steps
run() { python bench.py --no-extra --steps 5 --warmup 2 --realizations $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R', $1, 'chunk steps', '$2', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
for R in 128 148 256 296 444 512 600 768; do
for cs in 0 20; do SMM_ABM_CHUNK_STEPS=$cs run $R $cs; done
done
