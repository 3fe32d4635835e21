#!/bin/bash
# timing attribution of the dense ABM pair loop: variants with WRONG physics that drop one resource each
run() { python bench.py --no-extra --steps 6 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
run base
for v in ONE_MUFU NO_SHFL NO_ROWLDS; do SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_exp_$v.so run $v; done
run base
