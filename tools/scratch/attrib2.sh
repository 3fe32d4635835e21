#!/bin/bash
# with step chunks not every realization needs a resident CTA: fewer CTAs per SM with more registers each?
run() { python bench.py --no-extra --steps 6 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
run base7
for v in 6 5 4; do SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_exp_occ$v.so run occ$v; done
