#!/bin/bash
# wide CTAs: head start of warp 0 (which also updates influencers / media) in the tile dealing
run() { python bench.py --no-extra --steps 5 --warmup 2 --realizations $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R', $1, 'p2 cost', '$2', 'abm kernel ms', round(d['roofline']['kernel_ms'],4))"; }
for R in 128 296; do
for c in 0 10 20 30 45; do SMM_ABM_P2_COST=$c run $R $c; done
done
for c in 0 20 30 45; do SMM_ABM_P2_COST=$c python - <<'PY'
import os, sys; sys.path.insert(0, "/root/repo")
import socialmediamodel_b200 as smm
ctx = smm.Context.default()
smm.ABMsolve(100, chosenseed=0, ctx=ctx)
ms = []
for k in range(10):
    smm.ABMsolve(100, chosenseed=k, ctx=ctx); ms.append(ctx.kernel_ms(0))
print("C1 p2 cost", os.environ["SMM_ABM_P2_COST"], "kernel ms", round(min(ms), 4))
PY
done
