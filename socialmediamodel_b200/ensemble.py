"""ABMsolveensemble / PDEsolveensemble of the reference (src/ensemble.jl) on the B200 path,
plus the ensemble statistics of plotting.jl:361-363 (means) and :284-307 (follower numbers).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .abm import _network, _noise_struct
from .setting import ABMconstruct, parameters


def savetimes_abm(NT, savepoints):
    """ensemble.jl:21 (1-based indices into xs) converted to 0-based history indices."""
    return (np.round(np.linspace(1, NT, savepoints)).astype(np.int32) - 1).astype(np.int32)


def abm_ensemble_raw(NT, R, *, p, q, savesteps, sigma=0.1, first_realization=0, seed=0, precision="f32",
                     init_state=None, injected=None, network=None, want=("us", "zs", "ys", "sums", "numbers"),
                     pinned=False, ctx=None, us_dtype=np.float64):
    """smm_abm_ensemble for the local shard of R realizations; Julia-shaped outputs.  ``us_dtype=np.float32`` returns the
    densities as Float32 (smm_abm_ensemble_f32, FP32 mode only); leaving "us" out of ``want`` returns the sums only.
    A multi-GPU ``ctx`` (Context(devices=...)) shards the realizations over its GPUs inside the call."""
    ctx = _lib.Context.default() if ctx is None else ctx
    n, L, M = q.n, q.L, q.M
    Nx, Ny = p.X.shape
    savesteps = np.ascontiguousarray(savesteps, np.int32)
    sp_ = len(savesteps)
    alloc = (lambda shape, dt=np.float64: ctx.pinned_empty(shape, dt)) if pinned else \
            (lambda shape, dt=np.float64: np.empty(shape, dt, order="F"))
    out = {}
    us_dtype = np.dtype(us_dtype)
    assert us_dtype == np.float64 or (us_dtype == np.float32 and precision == "f32"), "Float32 densities need precision='f32'"
    if "us" in want: out["us"] = alloc((Nx, Ny, 2, L, sp_, R), us_dtype)
    if "zs" in want: out["zs"] = np.empty((2, 2, sp_, R), order="F")
    if "ys" in want: out["ys"] = np.empty((2, L, sp_, R), order="F")
    if "sums" in want:
        out["sum_us"] = np.empty((Nx, Ny, 2, L, sp_), order="F")
        out["sum_zs"] = np.empty((2, 2, sp_), order="F")
        out["sum_ys"] = np.empty((2, L, sp_), order="F")
    if "numbers" in want: out["numbers"] = np.empty((M, L, NT + 1), np.int64, order="F")
    keep = []
    if init_state is not None:
        x0 = _lib.f_array(np.asarray(init_state["x0"], float).reshape((n, 2, R), order="F"), np.float64)
        st = _lib.f_array(np.asarray(init_state["state"]).reshape((n, R), order="F"), np.int32)
        f0 = _lib.f_array(np.asarray(init_state["follow0"]).reshape((n, R), order="F"), np.int32)
        i0 = _lib.f_array(np.asarray(init_state["inf0"], float).reshape((L, 2, R), order="F"), np.float64)
        m0 = _lib.f_array(np.asarray(init_state["media0"], float).reshape((M, 2, R), order="F"), np.float64)
    else:
        x0 = st = f0 = i0 = m0 = None
    rowptr, colidx = _network(network, n)
    ns, keep = _noise_struct(seed, injected, n, L, M, NT, R)
    spar = _lib.make_params(q, p)
    g = _lib.make_grid(p)
    tail = (_lib.ptr(x0), _lib.ptr(st), _lib.ptr(f0), _lib.ptr(i0), _lib.ptr(m0),
            _lib.ptr(rowptr), _lib.ptr(colidx), C.byref(ns), sp_, _lib.ptr(savesteps),
            float(sigma), C.byref(g), _lib.ptr(out.get("us")), _lib.ptr(out.get("zs")),
            _lib.ptr(out.get("ys")), _lib.ptr(out.get("sum_us")), _lib.ptr(out.get("sum_zs")),
            _lib.ptr(out.get("sum_ys")), _lib.ptr(out.get("numbers")))
    if us_dtype == np.float32:
        rc = ctx.lib.smm_abm_ensemble_f32(ctx.h, C.byref(spar), int(NT), int(R), int(first_realization), *tail)
    else:
        rc = ctx.lib.smm_abm_ensemble(ctx.h, C.byref(spar), _lib.PREC[precision], int(NT), int(R), int(first_realization), *tail)
    ctx.check(rc, "smm_abm_ensemble")
    return out


def ABMsolveensemble(NT=200, N=10, *, savepoints=4, init="4inf", p=None, q=None, sigma=0.1, save=True,
                     precision="f32", ctx=None, devices=None, us_dtype=np.float64):
    """ensemble.jl:11-55: N realizations with seeds k = 1..N (ensemble.jl:25), KDE of the agents per
    (medium, influencer) at the save points.  Returns (us, zs, ys, (p, q)) with the reference's shapes
    us[N_x,N_y,2,L,savepoints,N], zs[2,2,savepoints,N], ys[2,L,savepoints,N].

    The reference seeds realization k with Random.seed!(k); here realization k uses the Philox
    stream (key = 0, realization = k), k = 1..N.  With save=True the arrays are written to
    data/abm_ensemble_<init>.npz under the reference's variable names (us2, zs2, ys2; ensemble.jl:46-53)."""
    p = ABMconstruct() if p is None else p
    q = parameters() if q is None else q
    own = None
    if devices is not None and ctx is None:   # keyword extension: several GPUs of the box inside this one call
        ctx = own = _lib.Context(devices=devices)
    try:
        out = abm_ensemble_raw(NT, N, p=p, q=q, savesteps=savetimes_abm(NT, savepoints), sigma=sigma,
                               first_realization=1, seed=0, precision=precision, want=("us", "zs", "ys"), ctx=ctx,
                               us_dtype=us_dtype)
    finally:
        if own is not None:
            own.close()
    us, zs, ys = out["us"], out["zs"], out["ys"]
    if save:
        os.makedirs("data", exist_ok=True)
        np.savez(os.path.join("data", f"abm_ensemble_{init}.npz"), us2=us, zs2=zs, ys2=ys)
    return us, zs, ys, (p, q)


def kde(x, follow, state, *, p, L, sigma=0.1, precision="f64", ctx=None):
    """smm_kde: us[Nx,Ny,2,L] = (1/n) sum of Gaussians per (medium, influencer) class (ensemble.jl:27-38)."""
    ctx = _lib.Context.default() if ctx is None else ctx
    n = len(follow)
    xx = _lib.f_array(np.asarray(x, float).reshape((n, 2), order="F"), np.float64)
    fo = np.ascontiguousarray(follow, np.int32); st = np.ascontiguousarray(state, np.int32)
    Nx, Ny = p.X.shape
    us = np.empty((Nx, Ny, 2, L), order="F")
    g = _lib.make_grid(p)
    ctx.check(ctx.lib.smm_kde(ctx.h, n, int(L), _lib.PREC[precision], _lib.ptr(xx), _lib.ptr(fo), _lib.ptr(st),
                              float(sigma), C.byref(g), _lib.ptr(us)), "smm_kde")
    return us


def runensembles(N, *, NT=200, tmax=2.0, savepoints=5, q=None, sigma=0.1, save=True, precision="f32"):
    """ensemble.jl:88-95 without the plotting call: one PDE solve and N ABM realizations saved at the same
    number of points.  Returns (us, zs, ys, (p, q), us2, zs2, ys2, (p2, q))."""
    q = parameters() if q is None else q
    us, zs, ys, (p, qp) = PDEsolveensemble(tmax, 1, savepoints=savepoints, q=q, save=save, precision=precision)
    us2, zs2, ys2, (p2, q2) = ABMsolveensemble(NT, N, savepoints=savepoints, q=q, sigma=sigma, save=save, precision=precision)
    return us, zs, ys, (p, qp), us2, zs2, ys2, (p2, qp)


def ensemble_means(us, zs, ys):
    """plotting.jl:361-363."""
    N = us.shape[-1]
    return us.sum(axis=-1) / N, zs.sum(axis=-1) / N, ys.sum(axis=-1) / N


def PDEsolveensemble(tmax=0.1, N=10, *, savepoints=4, alg=None, init="4inf", p=None, q=None, save=True, dt=None,
                     scheme="rk4", precision="f32", qs=None, rng=None, ctx=None, adaptive=False, abstol=1e-6, reltol=1e-3,
                     dtmin=0.0, force_dtmin=False):
    """ensemble.jl:58-85: N PDE solves saved at LinRange(0, tmax, savepoints).  The reference runs N
    *identical* solves unless init="random"; ``qs`` (a list of N parameter sets) turns the call into a
    parameter sweep.  All instances advance together as one device-resident batch.  Returns
    (us, zs, ys, (p, q)) with us[N_x,N_y,2,L,savepoints,N], zs[2,2,savepoints,N], ys[2,L,savepoints,N].
    ``adaptive=True`` (or ``alg="BS3"``): the adaptive integrator with per-instance step control and the reference's
    solve keywords instead of the fixed-step scheme."""
    from .pde import constructinitial, default_dt, pde_solve_adaptive_raw, pde_solve_raw
    from .setting import PDEconstruct
    p = PDEconstruct() if p is None else p
    q = parameters() if q is None else q
    inits = [constructinitial(init, (p, qq), rng) for qq in (qs if qs is not None else [q] * N)]
    qlist = [qq for _, qq in inits]
    u0 = np.stack([uz.x[0] for uz, _ in inits], axis=-1)
    z0 = np.stack([uz.x[1] for uz, _ in inits], axis=-1)
    y0 = np.stack([uz.x[2] for uz, _ in inits], axis=-1)
    if adaptive or (isinstance(alg, str) and alg.upper() in ("BS3", "ADAPTIVE")):
        times = np.linspace(0.0, tmax, savepoints) if savepoints > 1 else np.array([0.0])     # ensemble.jl:65
        out = pde_solve_adaptive_raw(p, qlist, u0, z0, y0, savetimes=times, abstol=abstol, reltol=reltol, dtmin=dtmin,
                                     force_dtmin=force_dtmin, precision=precision, ctx=ctx)
        us, zs, ys = out["us"], out["zs"], out["ys"]
        if save:
            os.makedirs("data", exist_ok=True)
            np.savez(os.path.join("data", f"pde_ensemble_{init}.npz"), us=us, zs=zs, ys=ys)
        return us, zs, ys, (p, qlist[-1])
    dt = min(default_dt(p, qq) for qq in qlist) if dt is None else dt
    per = max(1, int(np.ceil(tmax / (savepoints - 1) / dt - 1e-9))) if savepoints > 1 else max(1, int(np.ceil(tmax / dt - 1e-9)))
    nsteps = per * max(savepoints - 1, 1)
    dt = tmax / nsteps
    savesteps = (np.arange(savepoints) * per).astype(np.int32) if savepoints > 1 else np.array([0], np.int32)
    out = pde_solve_raw(p, qlist, u0, z0, y0, dt=dt, nsteps=nsteps, savesteps=savesteps, scheme=scheme,
                        precision=precision, ctx=ctx)
    us, zs, ys = out["us"], out["zs"], out["ys"]
    P = (p, qlist[-1])
    if save:
        os.makedirs("data", exist_ok=True)
        np.savez(os.path.join("data", f"pde_ensemble_{init}.npz"), us=us, zs=zs, ys=ys)
    return us, zs, ys, P
