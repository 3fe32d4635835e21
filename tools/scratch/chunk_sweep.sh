#!/bin/bash
# dense ABM kernel: steps per chunk (0 = one CTA per realization for the whole run); "head" = the previous build
run() { python bench.py --no-extra --steps 6 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms/pass', round(d['ms_per_step'],4), 'abm kernel ms', round(d['roofline']['kernel_ms'],4), 'mass', d['ensemble_mass_check'])"; }
[ -f socialmediamodel_b200/libsmm_b200_head.so ] && SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_head.so run head
for cs in 0 50 25 20 10 5; do
SMM_ABM_CHUNK_STEPS=$cs run "chunk steps $cs"
done
[ -f socialmediamodel_b200/libsmm_b200_head.so ] && SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_head.so run head
