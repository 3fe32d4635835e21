import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
import torch
p = smm.PDEconstruct(); q = smm.parameters()
uz, qq = smm.constructinitial("4inf", (p, q))
for cluster in (False, True, False, True):
    os.environ["SMM_PDE_CLUSTER"] = "1" if cluster else "0"
    for n in (200, 2000, 2000):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        smm.pde_solve_raw(p, [qq], uz.x[0][..., None], uz.x[1][..., None], uz.x[2][..., None], dt=0.001, nsteps=n, savesteps=[n], precision="f32")
        torch.cuda.synchronize(); t1 = time.perf_counter()
        dev = smm.Context.default().kernel_ms(4)
        print("cluster" if cluster else "chain  ", n, f"{(t1 - t0) * 1e3:.2f} ms total, {(t1 - t0) / n * 1e6:.1f} us per RK4 step (incl. create/readback); device {dev / n * 1e3:.2f} us per step", flush=True)
