#!/bin/bash
# dense ABM kernel time against CTAs per SM (128-thread CTAs, R = 148 k): latency-bound or throughput-bound?
for k in 1 2 3 4 5 6 7; do
R=$((148*k))
SMM_ABM_THREADS=128 python bench.py --no-extra --steps 5 --warmup 2 --realizations $R 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($k, $R, 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"
done
