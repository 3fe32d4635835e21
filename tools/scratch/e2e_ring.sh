#!/bin/bash
run() { echo -n "$* : "; env "$@" SMM_E2E_DEBUG=1 python bench.py --no-extra --steps 5 --warmup 3 2> gpurun_out/ring.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e ms', round(d['e2e']['ms_per_step'],2), end='  ')"; grep "widen pipeline" gpurun_out/ring.err | tail -2 | tr '\n' ' '; echo; }
for i in 1 2; do
run DEFAULT=1
run SMM_HOST_RING_MB=82
run SMM_HOST_RING_MB=130
run SMM_HOST_L3_MB=0
done
