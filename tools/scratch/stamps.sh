#!/bin/bash
# per-CTA start / end times of the dense ABM kernel (library built with -DSMM_DENSE_STAMPS), device pass only
SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_st.so python - <<'PY' 2>&1 | grep "^cta" > gpurun_out/stamps_ctas.txt
import sys; sys.path.insert(0, ".")
import bench, torch
import socialmediamodel_b200 as smm
ctx = smm.Context.default()
e = bench.AbmEnsemble(ctx, None, 1024, 1)
e.device_pass(1); ctx.sync(); torch.cuda.synchronize()
PY
wc -l gpurun_out/stamps_ctas.txt
