#!/bin/bash
# e2e of smm_abm_ensemble under different host-widen settings (tools/scratch)
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'])"; }
run SMM_HOST_WIDEN_SUB=1
run SMM_HOST_WIDEN_SUB=2
run SMM_HOST_WIDEN_SUB=4
run SMM_HOST_WIDEN_SUB=8
run SMM_HOST_WIDEN_SUB=4 SMM_HOST_THREADS=8
run SMM_HOST_WIDEN_SUB=4 SMM_HOST_THREADS=12
run SMM_HOST_WIDEN_SUB=4 SMM_HOST_THREADS=24
