"""Float64 restatement of the reference's fixed-target control evaluation (src/controlstrategies.jl:5-115) on
top of pde_oracle, with the same fixed-step RK4 the GPU path uses.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.  PARITY UNPINNED (no reference fixtures; the reference integrates adaptively)."""
import numpy as np

from . import pde_oracle as P


def _solve(u, z, y, T, savedt, dt, p, q, ci, cm):
    """Saved states at k*savedt <= T (saveat = 0:savedt:T, controlstrategies.jl:10,96) and the state at T itself
    (the restart value sol(T), :101): when T is not a multiple of savedt the remainder is integrated separately."""
    nsave = int(np.floor(T / savedt + 1e-9)); per = max(1, int(np.ceil(savedt / dt - 1e-9))); h = savedt / per
    saved = P.rk_solve(u, z, y, nsave * per, h, p, q, ci, cm, save_every=per) if nsave > 0 else [(0, u, z, y)]
    rem = T - nsave * savedt
    final = saved[-1][1:]
    if rem > 1e-12 * max(1.0, T):
        nrem = max(1, int(np.ceil(rem / h - 1e-9)))
        final = P.rk_solve(*final, nrem, rem / nrem, p, q, ci, cm)
    return saved, final  # saved: list of (step,u,z,y) incl. step 0


def prep(tequil, p, q, start="zero", savedt=0.05, dt=1e-3, countercontrol="no"):
    """controlstrategies.jl:5-63."""
    u0, z0, y0, ci, cm = P.uniforminit(p, q)
    u, z, y = _solve(u0, z0, y0, tequil, savedt, dt, p, q, ci, cm)[1]
    Lold = q["L"]
    if countercontrol == "no":
        L, ci2, cm2 = Lold + 1, np.append(ci, 3.0), cm
    elif countercontrol == "inf":
        L, ci2, cm2 = Lold + 2, np.append(ci, [1.0, 3.0]), cm
    else:
        L, ci2, cm2 = Lold + 1, np.append(ci, 3.0), np.array([0.0, 1.0])
    u2 = np.zeros(u.shape[:3] + (L,)); u2[..., :Lold] = u
    if start == "influencer":
        sl = y[:, 0].copy()
    elif start == "zero":
        sl = np.zeros(2)
    else:
        w = u2[:, :, 1, :].sum(axis=2)
        sl = np.array([(w * p["X"]).sum(), (w * p["Y"]).sum()]) / u2[:, :, 1, :].sum()
    y2 = np.zeros((2, L)); y2[:, :Lold] = y; y2[:, Lold:] = sl[:, None]
    slm = z[:, 1].copy() if countercontrol == "med" else None
    return (u2, z, y2), sl, dict(q, L=L), ci2, cm2, slm


def fixedtargets(ts, targets, startlocation, p, q, ci, cm, uzy0, savedt=0.05, dt=1e-3, countercontrol="no",
                 stubborntarget=None, startlocationmed=None):
    """controlstrategies.jl:67-115."""
    L = q["L"]
    if countercontrol != "no":
        q = dict(q, controltarget2=np.asarray(stubborntarget, float),
                 controlspeed2=np.linalg.norm(np.asarray(stubborntarget, float) - startlocation) / ts[-1])
    sl = np.asarray(startlocationmed if countercontrol == "med" else startlocation, float)
    u, z, y = uzy0
    fsum = 0.0; pen = 0.0
    for n in range(targets.shape[1]):
        target = targets[:, n]; Dt = ts[n + 1] - ts[n]
        speed = np.linalg.norm(target - sl) / Dt
        q1 = dict(q, controltarget1=target, controlspeed1=speed)
        saved, final = _solve(u, z, y, Dt, savedt, dt, p, q1, ci, cm)
        fsum += sum(s[1][:, :, :, L - 1].sum() / s[1].sum() for s in saved) * savedt
        pen += speed ** 2 * Dt
        sl = target
        u, z, y = final
    return fsum, np.sqrt(pen), (u, z, y)
