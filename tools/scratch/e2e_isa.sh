#!/bin/bash
run() { echo -n "$* : "; env "$@" python bench.py --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e ms', round(d['e2e']['ms_per_step'],2), 'widen GB/s', round(d['e2e']['floor']['host_widen_gbs'],1), 'floor', round(d['e2e']['floor']['floor_ms'],2))"; }
grep -o "avx512[a-z]*" /proc/cpuinfo | sort | uniq -c | head -5
run SMM_HOST_WIDEN_ISA=avx2
run SMM_HOST_WIDEN_ISA=avx512
run SMM_HOST_WIDEN_ISA=avx2
run SMM_HOST_WIDEN_ISA=avx512
