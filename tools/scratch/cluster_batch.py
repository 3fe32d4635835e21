"""Cluster-resident steps vs the kernel chain as a function of the batch size (where should the default switch?)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
ctx = smm.Context.default()
for name, p in (("coarse 47", smm.PDEconstructcoarse()), ("default 89", smm.PDEconstruct())):
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    for B in (1, 4, 9, 16, 27, 36, 64):
        row = []
        for cl in ("1", "0"):
            os.environ["SMM_PDE_CLUSTER"] = cl
            pb = smm.PDEBatch(p, [q2] * B, *uzy0.x, precision="f32", shared_init=True, ctx=ctx)
            pb.step(1e-3, 10); ctx.sync()
            ms = []
            for _ in range(3):
                pb.step(1e-3, 100); ctx.sync(); ms.append(ctx.kernel_ms(4) / 100)
            pb.close()
            row.append(float(np.median(ms)) * 1e3)
        print(f"{name} B={B:3d}  cluster {row[0]:7.1f} us/step   chain {row[1]:7.1f} us/step", flush=True)
