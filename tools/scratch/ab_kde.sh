#!/bin/bash
run() { python bench.py --no-extra --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms/pass', round(d['ms_per_step'],4), 'kde ms', d.get('roofline_kde',{}).get('kernel_ms'), 'mass', d['ensemble_mass_check'])"; }
run kc32
for kc in 48 64; do SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_kc$kc.so run kc$kc; done
run kc32
