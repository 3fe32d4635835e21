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
def solve(prec, n=8):
    return smm.pde_solve_raw(p, ql, U, Z, Y, dt=5e-5, nsteps=n, savesteps=[n], precision=prec)["us"]
ref = solve("f64")
sc = np.abs(ref).max()
runs = [solve("f32") for _ in range(6)]
print(os.environ.get("TAG"), "err vs f64:", [f"{np.abs(r - ref).max() / sc:.2e}" for r in runs], " vs run0:", [f"{np.abs(r - runs[0]).max() / sc:.1e}" for r in runs[1:]])
