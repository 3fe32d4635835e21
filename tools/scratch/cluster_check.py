"""Cluster-resident PDE steps (pde_cluster.cu) against the kernel chain and the Float64 chain; timing of both."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm


def run(p, ql, U, Z, Y, nsteps, scheme, prec, cluster):
    os.environ["SMM_PDE_CLUSTER"] = "1" if cluster else "0"
    return smm.pde_solve_raw(p, ql, U, Z, Y, dt=2e-3 if p.N_x < 60 else 1e-3, nsteps=nsteps, savesteps=[nsteps // 2, nsteps], scheme=scheme, precision=prec)


for name, p in (("coarse 47", smm.PDEconstructcoarse()), ("default 89", smm.PDEconstruct())):
    for B in (1, 3):
        qs = [smm.parameters(), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)][:B]
        inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
        ql = [qq for _, qq in inits]
        rng = np.random.default_rng(3)
        xx = np.linspace(0, 3, p.N_x)[:, None, None, None]; yy = np.linspace(0, 2, p.N_y)[None, :, None, None]
        U = np.stack([uz.x[0] * (1 + 0.2 * np.sin(xx * (b + 1)) * np.cos(yy)) for b, (uz, _) in enumerate(inits)], axis=-1)
        Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
        for scheme in ("rk4", "euler", "heun", "ssprk3"):
            ref = run(p, ql, U, Z, Y, 20, scheme, "f64", False)
            a = run(p, ql, U, Z, Y, 20, scheme, "f32", False)
            c = run(p, ql, U, Z, Y, 20, scheme, "f32", True)
            sc = np.abs(ref["us"]).max()
            print(f"{name} B={B} {scheme:7s} chain-f64 {np.abs(a['us'] - ref['us']).max() / sc:.2e}  cluster-f64 {np.abs(c['us'] - ref['us']).max() / sc:.2e}"
                  f"  ys: chain {np.abs(a['ys'] - ref['ys']).max():.2e} cluster {np.abs(c['ys'] - ref['ys']).max():.2e}", flush=True)

# timing: C3, one instance, 1000 RK4 steps
import torch
p = smm.PDEconstruct(); q = smm.parameters()
uz, qq = smm.constructinitial("4inf", (p, q))
for cluster in (False, True):
    os.environ["SMM_PDE_CLUSTER"] = "1" if cluster else "0"
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        smm.pde_solve_raw(p, [qq], uz.x[0][..., None], uz.x[1][..., None], uz.x[2][..., None], dt=0.001, nsteps=2000, savesteps=[2000], precision="f32")
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print("cluster" if cluster else "chain", f"{(t1 - t0) / 2000 * 1e6:.1f} us per RK4 step (incl. create/readback)")
