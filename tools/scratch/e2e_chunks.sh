#!/bin/bash
run() { echo -n "$* : "; env "$@" python bench.py --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e ms', round(d['e2e']['ms_per_step'],2), 'steps', d['e2e']['steps'], 'floor', round(d['e2e']['floor']['floor_ms'],2))"; }
run SMM_ENSEMBLE_CHUNKS=4 SMM_HOST_WIDEN_SUB=2
run SMM_ENSEMBLE_CHUNKS=8 SMM_HOST_WIDEN_SUB=4
run SMM_ENSEMBLE_CHUNKS=4 SMM_HOST_WIDEN_SUB=2
run SMM_ENSEMBLE_CHUNKS=8 SMM_HOST_WIDEN_SUB=4
run SMM_ENSEMBLE_CHUNKS=8 SMM_HOST_WIDEN_SUB=3
run SMM_ENSEMBLE_CHUNKS=10 SMM_HOST_WIDEN_SUB=3
