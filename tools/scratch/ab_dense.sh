#!/bin/bash
# A/B timing of a dense-ABM kernel variant built as its own library (SMM_B200_LIB)
for i in 1 2; do
python bench.py --no-extra --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base', d['ms_per_step'], d['roofline']['kernel_ms'])"
SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_vee.so python bench.py --no-extra --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
