#!/bin/bash
for cfg in "8 4" "8 8" "4 16" "8 6"; do set -- $cfg
echo "== chunks $1 sub $2"
SMM_E2E_DEBUG=1 SMM_ENSEMBLE_CHUNKS=$1 SMM_HOST_WIDEN_SUB=$2 python bench.py --no-extra --steps 3 --warmup 2 2>&1 >/dev/null | grep "widen pipeline" | tail -3
done
