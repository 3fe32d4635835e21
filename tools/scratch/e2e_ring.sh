#!/bin/bash
run() { echo -n "$* : "; env "$@" SMM_E2E_DEBUG=1 python bench.py --no-extra --steps 5 --warmup 3 2> gpurun_out/ring.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e ms', round(d['e2e']['ms_per_step'],2), end='  ')"; grep "widen pipeline" gpurun_out/ring.err | tail -1; }
run DEFAULT=1
run SMM_HOST_WIDEN_SUB=3
run SMM_HOST_WIDEN_SUB=5
run DEFAULT=1
run SMM_HOST_WIDEN_SUB=3
run SMM_HOST_L3_MB=0
