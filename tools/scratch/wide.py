import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
p, q = smm.ABMconstruct(), smm.parameters()
st = smm.savetimes_abm(200, 4)
ctx = smm.Context.default()
ref = None
for thr in ("512", "512"):
    os.environ["SMM_ABM_THREADS"] = thr
    for R in (128, 1):
        ms = []
        for _ in range(5):
            a = smm.abm_ensemble_raw(200, R, p=p, q=q, savesteps=st, first_realization=1, precision="f32", want=("ys", "sums"))
            ms.append(ctx.kernel_ms(0))
        if ref is None: ref = a["ys"].copy()
        print(thr, "R", R, "dense kernel ms", f"{np.median(ms):.4f}", "us per step", f"{np.median(ms) / 200 * 1e3:.2f}", "same as 512:", np.array_equal(a["ys"], ref) if R == 128 else "-")
