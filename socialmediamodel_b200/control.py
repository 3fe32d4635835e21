"""Control layer of the reference as a caller of the GPU right-hand side (src/controlstrategies.jl:5-132;
SURVEY section 8(f) "next" #1): prep, PDEsolvefixedtargetsfast, PDEsolvefixedtargets.

The optimiser (NLopt COBYLA/MLSL, controlstrategies.jl:137-209) is out of scope; what it evaluates -- the
objective (follower share integrated over time, L2 norm of the control speed) for given targets -- is here,
and ``fixedtargets_objectives`` evaluates many candidate strategies at once as ONE device-resident batch
(the ensemble axis of smm_pde_solve), which is what a multistart search needs.

Time integration: the reference uses the adaptive default solver with abstol / reltol / dtmin / force_dtmin and
saveat = 0:savedt:T (controlstrategies.jl:10,96).  ``integrator="adaptive"`` maps those keywords onto the
device-side embedded 3(2) integrator (smm_pde_solve_adaptive; every candidate of a batch has its own step
control); the default ``integrator="fixed"`` is the fixed-step RK4 graph with a step that divides ``savedt``
(bitwise reproducible against the oracle).  Reference quirks kept: with
countercontrol="no" the added influencer carries flag 3 and therefore moves towards controltarget2 at
controlspeed2 (pde.jl:249-255), not towards ``targets`` (controlstrategies.jl:16,92).
"""
import numpy as np

from .pde import ArrayPartition, PDESolution, constructinitial, default_dt, pde_solve_adaptive_raw, pde_solve_raw
from .setting import PDEconstruct, PDEconstructcoarse, parameterscontrol, parametersstronginf


def _segment_steps(T, savedt, dt):
    nsave = int(np.floor(T / savedt + 1e-9))                 # saveat = 0:savedt:T keeps the points k*savedt <= T
    per = max(1, int(np.ceil(savedt / dt - 1e-9)))
    return nsave, per, savedt / per


def _solve_segments(p, qs, u0, z0, y0, T, savedt, dt, precision, ctx, adapt=None):
    """Integrate B instances over [0,T].  Returns (out, times, final): the states saved every savedt (like
    saveat=0:savedt:T, controlstrategies.jl:10,96) and ``final`` = (u, z, y) at T itself, the restart value
    sol(T) (:101).  When T is not a multiple of savedt (the reference's defaults: Tmax/ntarg = 5/3) the reference
    evaluates sol(T) on a saveat-only solution, i.e. by extrapolating the last saved interval; here the remainder
    is integrated.  adapt = None: fixed step; else dict(atol, rtol, dtmin) for the adaptive integrator
    (force_dtmin = true)."""
    nsave, per, h = _segment_steps(T, savedt, dt)
    rem = T - nsave * savedt
    has_rem = rem > 1e-12 * max(1.0, T)
    if adapt is not None:
        times = np.arange(nsave + 1) * savedt
        allt = np.append(times, T) if has_rem else times
        out = pde_solve_adaptive_raw(p, qs, u0, z0, y0, savetimes=allt, abstol=adapt["atol"], reltol=adapt["rtol"],
                                     dtmin=adapt["dtmin"], force_dtmin=True, precision=precision, ctx=ctx)
        final = (out["us"][..., -1, :], out["zs"][..., -1, :], out["ys"][..., -1, :])
        if has_rem:
            out = dict(out, us=out["us"][..., :-1, :], zs=out["zs"][..., :-1, :], ys=out["ys"][..., :-1, :])
        return out, times, final
    savesteps = (np.arange(nsave + 1) * per).astype(np.int32)
    out = pde_solve_raw(p, qs, u0, z0, y0, dt=h, nsteps=nsave * per, savesteps=savesteps, precision=precision, ctx=ctx)
    final = (out["us"][..., -1, :], out["zs"][..., -1, :], out["ys"][..., -1, :])
    if has_rem:
        nrem = max(1, int(np.ceil(rem / h - 1e-9)))
        tail = pde_solve_raw(p, qs, *final, dt=rem / nrem, nsteps=nrem, savesteps=np.array([nrem], np.int32),
                             precision=precision, ctx=ctx)
        final = (tail["us"][..., 0, :], tail["zs"][..., 0, :], tail["ys"][..., 0, :])
    return out, savesteps * h, final


def _adapt(integrator, atol, rtol, dtmin):
    if integrator == "fixed":
        return None
    if integrator != "adaptive":
        raise ValueError(f"unknown integrator {integrator!r}")
    return dict(atol=atol, rtol=rtol, dtmin=dtmin)


def prep(tequil=5.0, *, p=None, q=None, r=None, init="uniform", savedt=0.05, atol=1e-6, rtol=1e-3, dtmin=0.001,
         countercontrol="no", returnsol=False, dt=None, precision="f32", ctx=None, integrator="fixed"):
    """controlstrategies.jl:5-63: equilibrate for tequil, then append the controlled influencer(s)."""
    p = PDEconstruct() if p is None else p
    q = parametersstronginf() if q is None else q
    r = parameterscontrol() if r is None else r
    uzy0, q1 = constructinitial(init, (p, q))
    dt = default_dt(p, q1) if dt is None else dt
    out, ts, final = _solve_segments(p, [q1], *uzy0.x, tequil, savedt, dt, precision, ctx, _adapt(integrator, atol, rtol, dtmin))
    sol1 = PDESolution(ts, [ArrayPartition(out["us"][..., m, 0], out["zs"][..., m, 0], out["ys"][..., m, 0]) for m in range(len(ts))])
    u, z, y = final[0][..., 0], final[1][..., 0], final[2][..., 0]
    Lold = q1.L
    ci = list(np.asarray(q1.controlled_inf, float))
    if countercontrol == "no":
        q2 = q1.merge(L=Lold + 1, controlled_inf=np.array(ci + [3.0]))
    elif countercontrol == "inf":
        q2 = q1.merge(L=Lold + 2, controlled_inf=np.array(ci + [1.0, 3.0]))
    elif countercontrol == "med":
        q2 = q1.merge(L=Lold + 1, controlled_inf=np.array(ci + [3.0]), controlled_med=np.array([0.0, 1.0]))
    else:
        raise ValueError(f"unknown countercontrol {countercontrol!r}")
    L = q2.L
    u2 = np.zeros((p.N_x, p.N_y, 2, L), order="F")
    u2[:, :, :, :Lold] = u
    if r.start == "influencer":
        startlocation = y[:, 0].reshape(1, 2).copy()
    elif r.start == "zero":
        startlocation = np.zeros((1, 2))
    elif r.start == "mean":   # mean opinion of the agents of medium 2 (controlstrategies.jl:36)
        w = u2[:, :, 1, :].sum(axis=2)
        startlocation = (np.array([(w * p.X).sum(), (w * p.Y).sum()]) / u2[:, :, 1, :].sum()).reshape(1, 2)
    else:
        raise ValueError(f"unknown start {r.start!r}")
    y2 = np.zeros((2, L), order="F")
    y2[:, :Lold] = y
    y2[:, Lold:] = startlocation.reshape(2, 1)
    startlocationmed = z[:, 1].reshape(1, 2).copy() if countercontrol == "med" else None
    uzy = ArrayPartition(u2, np.asfortranarray(z), y2)
    if not returnsol:
        return uzy, startlocation, (p, q2), startlocationmed
    return uzy, startlocation, (p, q2), startlocationmed, sol1, (p, q1)


def fixedtargets_objectives(ts, targets_list, startlocation, pq, r, uzy0, *, savedt=0.05, atol=1e-6, rtol=1e-3, dtmin=0.001,
                            countercontrol="no", stubborntarget=None, startlocationmed=None, returnsol=False, dt=None,
                            precision="f32", ctx=None, integrator="fixed"):
    """PDEsolvefixedtargetsfast (controlstrategies.jl:67-115) for K candidate strategies at once: targets_list[k]
    is a (2, n_targets) array.  Returns followersum[K], speedpenalty[K] (and the per-candidate sols/Ps)."""
    p, q = pq
    K = len(targets_list)
    ts = np.ravel(np.asarray(ts, float))
    L = q.L
    if countercontrol != "no":
        stub = np.asarray(stubborntarget, float).reshape(1, 2)
        q = q.merge(controltarget2=stub, controlspeed2=float(np.linalg.norm(stub - startlocation) / ts[-1]))
    start = [np.asarray(startlocationmed if countercontrol == "med" else startlocation, float).reshape(1, 2)] * K
    targets_list = [np.asarray(t, float).reshape(2, -1) for t in targets_list]
    n_targets = targets_list[0].shape[1]
    dt = default_dt(p, q) if dt is None else dt
    followersum = np.zeros(K); speedpen = np.zeros(K)
    u = np.stack([uzy0.x[0]] * K, axis=-1); z = np.stack([uzy0.x[1]] * K, axis=-1); y = np.stack([uzy0.x[2]] * K, axis=-1)
    sols = [[] for _ in range(K)]; Ps = [[] for _ in range(K)]
    for n in range(n_targets):
        Dt = ts[n + 1] - ts[n]
        qs = []
        for k in range(K):
            target = targets_list[k][:, n].reshape(1, 2)
            speed = float(np.linalg.norm(target - start[k]) / Dt)
            qs.append(q.merge(controltarget1=target, controlspeed1=speed))
            speedpen[k] += speed ** 2 * Dt
            start[k] = target
        out, tsave, final = _solve_segments(p, qs, u, z, y, Dt, savedt, dt, precision, ctx, _adapt(integrator, atol, rtol, dtmin))
        us = out["us"]                                                       # [Nx,Ny,2,L,nsave,K]
        share = us[:, :, :, L - 1].sum(axis=(0, 1, 2)) / us.sum(axis=(0, 1, 2, 3))   # [nsave, K]
        followersum += share.sum(axis=0) * savedt                            # controlstrategies.jl:98
        u, z, y = final
        if returnsol:
            for k in range(K):
                sols[k].append(PDESolution(tsave, [ArrayPartition(us[..., m, k], out["zs"][..., m, k], out["ys"][..., m, k])
                                                   for m in range(len(tsave))]))
                Ps[k].append((p, qs[k]))
    speedpen = np.sqrt(speedpen)
    return (followersum, speedpen, sols, Ps) if returnsol else (followersum, speedpen)


def PDEsolvefixedtargetsfast(ts, targets, startlocation, pq, r, uzy0, *, savedt=0.05, atol=1e-6, rtol=1e-3, dtmin=0.001,
                             countercontrol="no", stubborntarget=None, startlocationmed=None, returnsol=False, dt=None,
                             precision="f32", ctx=None, integrator="fixed"):
    """controlstrategies.jl:67-115 (atol / rtol / dtmin are used by integrator="adaptive")."""
    res = fixedtargets_objectives(ts, [targets], startlocation, pq, r, uzy0, savedt=savedt, atol=atol, rtol=rtol, dtmin=dtmin,
                                  countercontrol=countercontrol, stubborntarget=stubborntarget,
                                  startlocationmed=startlocationmed, returnsol=returnsol, dt=dt, precision=precision, ctx=ctx,
                                  integrator=integrator)
    if returnsol:
        return float(res[0][0]), float(res[1][0]), res[2][0], res[3][0]
    return float(res[0][0]), float(res[1][0])


def PDEsolvefixedtargets(targets, *, p=None, q=None, r=None, init="uniform", countercontrol="no", stubborntarget=(1.5, 1.5),
                         savedt=0.05, dt=None, precision="f32", ctx=None, integrator="fixed"):
    """controlstrategies.jl:120-132."""
    p = PDEconstructcoarse() if p is None else p
    q = parametersstronginf() if q is None else q
    r = parameterscontrol(ntarg=1, start="zero") if r is None else r
    ntarg, Tmax, tequil = r.ntarg, r.Tmax, r.tequil
    dtmin = r.get("dtmin", 0.001)
    ts = np.concatenate([[0.0], Tmax / ntarg * np.arange(1, ntarg + 1)])
    targets = np.asarray(targets, float).reshape((2, ntarg), order="F")
    uzy0, startlocation, (p, q2), slm, sol1, (p, q1) = prep(tequil, p=p, q=q, r=r, init=init, savedt=savedt,
                                                            countercontrol=countercontrol, returnsol=True, dt=dt,
                                                            precision=precision, ctx=ctx, integrator=integrator, dtmin=dtmin)
    fs, sp, sols, Ps = PDEsolvefixedtargetsfast(ts, targets, startlocation, (p, q2), r, uzy0, savedt=savedt,
                                                countercontrol=countercontrol,
                                                stubborntarget=np.asarray(stubborntarget, float).reshape(1, 2),
                                                startlocationmed=slm, returnsol=True, dt=dt, precision=precision, ctx=ctx,
                                                integrator=integrator, dtmin=dtmin)
    return fs, sp, [sol1] + sols, [(p, q1)] + Ps


def PDEsolveopt(*, p=None, q=None, r=None, alg="LN_COBYLA", mtime=1000, meval=-1, ftol_rel=1e-4, xtol_rel=1e-2, x0=None,
                countercontrol="no", stubborntarget=None, multistart=False, save=True, savedt=0.05, dt=None, precision="f32",
                ctx=None, integrator="fixed", nstarts=8, rng=None):
    """controlstrategies.jl:137-209: maximise the follower objective over the 2*ntarg target coordinates.

    The reference hands the objective to NLopt (COBYLA, optionally MLSL multistart), one evaluation at a time.
    NLopt is not part of the reference tree and is not reproduced; what is kept is the problem (objective incl.
    its quirk that alpha multiplies the speed penalty twice for maximize="follower", `:156,165`; box bounds =
    the domain, `:186-189`; stopping rules mtime / meval / ftol_rel / xtol_rel, `:194-198`; return tuple `:206`)
    and it is solved with a derivative-free compass search whose 2*ndim polling points per iteration are
    evaluated as ONE device batch (`fixedtargets_objectives`).  multistart: `nstarts` low-discrepancy starting
    points (the role of G_MLSL_LDS) are evaluated as one batch and the search starts from the best.
    ``alg`` is accepted for signature compatibility.  Returns (maxf, maxx, ret, numevals, x_list,
    followersum_list, penalty_list)."""
    import time as _time
    p = PDEconstructcoarse() if p is None else p
    q = parametersstronginf() if q is None else q
    r = parameterscontrol() if r is None else r
    ntarg, Tmax, tequil, dtmin, maximize, alpha = r.ntarg, r.Tmax, r.tequil, r.dtmin, r.maximize, r.alpha
    if maximize not in ("follower", "counter_follower"):
        raise ValueError(f"This obj is unknown: {maximize!r}")
    uzy0, startlocation, (p, q2), slm = prep(tequil, p=p, q=q, r=r, dtmin=dtmin, countercontrol=countercontrol, savedt=savedt,
                                             dt=dt, precision=precision, ctx=ctx, integrator=integrator)
    ts = np.concatenate([[0.0], Tmax / ntarg * np.arange(1, ntarg + 1)])
    ndim = 2 * ntarg
    lb, ub = float(p.domain[0, 0]), float(p.domain[0, 1])
    x_list, followersum_list, penalty_list = [], [], []
    t_start = _time.perf_counter()

    def objective(X):
        """objective of every row of X (one device batch); records the evaluations like Objfun (:148-171)"""
        cands = [np.reshape(x, (2, ntarg), order="F") for x in X]
        fs, sp = fixedtargets_objectives(ts, cands, startlocation, (p, q2), r, uzy0, savedt=savedt, dtmin=dtmin,
                                         countercontrol=countercontrol, stubborntarget=stubborntarget, startlocationmed=slm,
                                         dt=dt, precision=precision, ctx=ctx, integrator=integrator)
        pen = alpha * sp
        for x, f_, s_ in zip(X, fs, pen):
            x_list.append(np.array(x)); followersum_list.append(float(f_)); penalty_list.append(float(s_))
        return fs - alpha * pen if maximize == "follower" else -fs - pen

    def budget_left(n):
        if meval > 0 and len(x_list) + n > meval:
            return False
        return _time.perf_counter() - t_start < mtime

    x = np.zeros(ndim) if x0 is None else np.clip(np.asarray(x0, float).ravel(), lb, ub)
    if multistart:
        try:
            from scipy.stats import qmc
            pts = qmc.scale(qmc.Sobol(ndim, seed=0 if rng is None else rng).random(nstarts), [lb] * ndim, [ub] * ndim)
        except ImportError:
            pts = np.random.default_rng(0 if rng is None else rng).uniform(lb, ub, (nstarts, ndim))
        starts = np.vstack([x[None, :], pts])
        fstart = objective(starts)
        x, fx = starts[int(np.argmax(fstart))].copy(), float(np.max(fstart))
    else:
        fx = float(objective(x[None, :])[0])
    step, ret = 0.25 * (ub - lb), "MAXEVAL_REACHED"
    while True:
        if step < xtol_rel * max(np.abs(x).max(), 1e-3 * (ub - lb)):
            ret = "XTOL_REACHED"
            break
        poll = np.repeat(x[None, :], 2 * ndim, axis=0)
        for i in range(ndim):
            poll[2 * i, i] = min(ub, x[i] + step)
            poll[2 * i + 1, i] = max(lb, x[i] - step)
        poll = poll[np.any(poll != x[None, :], axis=1)]
        if not budget_left(len(poll)):
            ret = "MAXEVAL_REACHED" if (meval > 0 and len(x_list) + len(poll) > meval) else "MAXTIME_REACHED"
            break
        fp = objective(poll)
        k = int(np.argmax(fp))
        if fp[k] > fx:
            gain = fp[k] - fx
            x, fx = poll[k].copy(), float(fp[k])
            if gain < ftol_rel * abs(fx):
                ret = "FTOL_REACHED"
                break
        else:
            step *= 0.5
    numevals = len(x_list)
    if save:
        import os
        os.makedirs("data", exist_ok=True)
        name = f"data/opt_strategy{ntarg}{alg}{maximize}{r.start}{str(bool(multistart)).lower()}{countercontrol}.npz"   # :203-205
        np.savez(name, maxf=fx, maxx=x, ret=ret, numevals=numevals, x_list=np.array(x_list),
                 followersum_list=np.array(followersum_list), penalty_list=np.array(penalty_list))
    return fx, x, ret, numevals, x_list, followersum_list, penalty_list
