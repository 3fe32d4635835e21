"""FP32 vs Float64 trajectory drift of the PDE path over long horizons (hard part 5 of SURVEY section 7)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import socialmediamodel_b200 as smm

for name, p, dt, nsteps in (("47", smm.PDEconstructcoarse(), 1e-3, 2000), ("89", smm.PDEconstruct(), 1e-3, 2000),
                            ("177", smm.PDEconstruct(dx=0.025), 5e-4, 1000)):
    q = smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    save = [nsteps // 20, nsteps // 4, nsteps]
    o32 = smm.pde_solve_raw(p, [q2], *uzy0.x, dt=dt, nsteps=nsteps, savesteps=save, precision="f32")
    o64 = smm.pde_solve_raw(p, [q2], *uzy0.x, dt=dt, nsteps=nsteps, savesteps=save, precision="f64")
    for m, k in enumerate(save):
        a, b = o32["us"][..., m, 0], o64["us"][..., m, 0]
        print(name, "step", k, "t=%.3f" % (k * dt), "rel err u %.2e" % (np.abs(a - b).max() / np.abs(b).max()),
              "y err %.2e" % np.abs(o32["ys"][..., m, 0] - o64["ys"][..., m, 0]).max(), "mass %.4f" % (b.sum() * p.dV),
              "min u %.2e" % b.min())
