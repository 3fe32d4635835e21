import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import socialmediamodel_b200 as smm
os.environ.setdefault("SMM_FFT_FAST", "1"); os.environ.setdefault("SMM_PDE_STAGE2", "1")
p = smm.PDEconstruct(dx=0.0125)
qs = [smm.parameters(a=0.5, sigma=0.75), smm.parameters(a=2.0, b=6.0, eta=1.0), smm.parameters(c=4.0, sigma=0.25)]
inits = [smm.constructinitial("4inf", (p, q)) for q in qs]
ql = [qq for _, qq in inits]
rng = np.random.default_rng(3)
U = np.stack([uz.x[0] * (1 + 0.2 * rng.random(uz.x[0].shape)) for uz, _ in inits], axis=-1)
Z = np.stack([uz.x[1] for uz, _ in inits], axis=-1); Y = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
def solve(scheme):
    return smm.pde_solve_raw(p, ql, U, Z, Y, dt=5e-5, nsteps=6, savesteps=list(range(1, 7)), scheme=scheme, precision="f32")
for scheme in ("euler", "rk4"):
    runs = [solve(scheme) for _ in range(4)]
    for r in runs[1:]:
        d = np.abs(r["us"] - runs[0]["us"])          # [Nx,Ny,2,L,ns,B]
        print(scheme, "per step max diff:", [float(d[..., k, :].max()) for k in range(6)], " ys:", float(np.abs(r["ys"] - runs[0]["ys"]).max()))
        k = int(np.argmax([d[..., kk, :].max() > 0 for kk in range(6)]))
        if d[..., k, :].max() > 0:
            idx = np.argwhere(d[..., k, :] > 0)
            print("   first step", k + 1, "mismatches", len(idx), "x range", idx[:, 0].min(), idx[:, 0].max(), "y range", idx[:, 1].min(), idx[:, 1].max(),
                  "planes", sorted(set(map(tuple, idx[:, 2:4].tolist())))[:8], "inst", sorted(set(idx[:, 4].tolist())))
