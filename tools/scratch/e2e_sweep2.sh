#!/bin/bash
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'])"; }
run SMM_HOST_WIDEN_SUB=2
run SMM_HOST_WIDEN_SUB=16
run SMM_HOST_WIDEN_SUB=32
run SMM_HOST_WIDEN_SUB=64
run SMM_HOST_WIDEN_SUB=2
