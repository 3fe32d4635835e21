"""PDE side of the reference (src/pde.jl) on the B200 path: the right-hand side ``f``, PDEsolve,
the initial conditions, and a device-resident batch for ensembles / parameter sweeps.

The reference hands ``f`` to DifferentialEquations.solve (pde.jl:290-291; adaptive, un-vendored).
Here time integration is a fixed-step explicit Runge-Kutta scheme whose whole stage chain is
captured in a CUDA graph (smm_pde_batch_step); ``PDEsolve`` keeps the reference's signature and
returns a callable ``sol`` with ``sol.t`` and ``sol(t).x == (u, z, y)`` (pde.jl:277-282).
"""
import ctypes as C

import numpy as np

from . import _lib
from .setting import PDEconstruct, Struct, parameters


class ArrayPartition:
    """RecursiveArrayTools.ArrayPartition stand-in: ``.x`` is the tuple (u, z, y) (pde.jl:118,189)."""

    def __init__(self, *parts):
        self.x = tuple(parts)


def _round_half_even(v):
    return int(np.round(v))


def uniforminit(pq):
    """pde.jl:90-119."""
    p, q = pq
    assert q.M == 2 and q.L == 4
    N_x, N_y, dV, domain, dx = p.N_x, p.N_y, p.dV, p.domain, p.dx
    rho = np.zeros((N_x, N_y, 2, 4), order="F")
    mid_y, mid_x = _round_half_even(N_y / 2), _round_half_even(N_x / 2)
    start_x = _round_half_even((domain[0, 1] - 2) / dx + 1)
    end_x = N_x - start_x + 1
    s = start_x - 1
    rho[s:mid_x, s:mid_y, :, 3] = 1
    rho[s:mid_x, mid_y:end_x, :, 1] = 1
    rho[mid_x:end_x, s:mid_y, :, 2] = 1
    rho[mid_x:end_x, mid_y:end_x, :, 0] = 1
    u0 = rho / (rho.sum() * dV)
    z0 = np.array([[-1.0, 1.0], [-1.0, 1.0]], order="F")
    y0 = np.array([[1.0, -1.0, 1.0, -1.0], [1.0, 1.0, -1.0, -1.0]], order="F")
    return ArrayPartition(u0, z0, y0), np.zeros(4), np.zeros(2)


def gaussian(x, center, sigma=0.1):
    """pde.jl:122."""
    d = np.asarray(x) - np.asarray(center)
    return 1 / (2 * np.pi * sigma ** 2) * np.exp(-1 / (2 * sigma ** 2) * (d ** 2).sum(axis=-1))


def randominit(pq, rng=None):
    """pde.jl:124-168: KDE of n uniformly sampled agents (the reference draws from the unseeded
    global RNG; pass ``rng`` for reproducibility)."""
    p, q = pq
    rng = np.random.default_rng() if rng is None else rng
    n, L = q.n, q.L
    assert q.M == 2 and L == 4
    dom = p.domain
    pos = rng.random((n, 2)) * (dom[0, 1] - dom[0, 0]) + dom[0, 0]
    med = rng.integers(0, 2, n)
    inf = np.where(pos[:, 1] > 0, np.where(pos[:, 0] > 0, 0, 1), np.where(pos[:, 0] > 0, 2, 3))
    # density = KDE (sigma = 0.1, pde.jl:122,136) of the sampled agents on the PDE grid: the GPU kernel K4
    from .ensemble import kde
    rho = kde(pos, inf, np.where(med == 1, 1, -1), p=p, L=L, sigma=0.1, precision="f64") * n
    u0 = rho / (rho.sum() * p.dV)
    counts = np.array([[np.sum((inf == j) & (med == i)) for i in range(2)] for j in range(L)], float)
    y0 = np.stack([[pos[inf == j, d].sum() for j in range(L)] for d in range(2)]) / counts.sum(axis=1)[None, :]
    z0 = np.array([[-1.0, 1.0], [-1.0, 1.0]], order="F")
    return ArrayPartition(u0, z0, np.asfortranarray(y0)), np.zeros(L), np.zeros(2)


def constructinitial(init, pq, rng=None):
    """pde.jl:170-180."""
    p, q = pq
    if init in ("4inf", "uniform"):
        uzy0, ci, cm = uniforminit((p, q))
    elif init == "random":
        uzy0, ci, cm = randominit((p, q), rng)
    else:
        raise ValueError(f"unknown init {init!r}")
    return uzy0, q.merge(controlled_inf=ci, controlled_med=cm)


def _params_array(qs, p):
    arr = (_lib.smm_params * len(qs))()
    for k, q in enumerate(qs):
        arr[k] = _lib.make_params(q, p)
    return arr


def _controls(qs, L):
    ci = np.zeros((L, len(qs)), order="F"); cm = np.zeros((2, len(qs)), order="F")
    for k, q in enumerate(qs):
        if "controlled_inf" in q: ci[:, k] = np.asarray(q.controlled_inf, float)
        if "controlled_med" in q: cm[:, k] = np.asarray(q.controlled_med, float)
    return ci, cm


def pde_rhs(u, z, y, *, p, q, precision="f32", ctx=None):
    """du, dz, dy = f(u, z, y) for one instance (u[Nx,Ny,2,L]) or a batch (u[Nx,Ny,2,L,B] with a list
    of q).  Backs the reference's f(duzy, uzy, (p,q), t) (pde.jl:183-268)."""
    ctx = _lib.Context.default() if ctx is None else ctx
    single = u.ndim == 4
    qs = [q] if isinstance(q, dict) else list(q)
    B = len(qs)
    L = u.shape[3]
    Nx, Ny = u.shape[0], u.shape[1]
    uu = _lib.f_array(np.asarray(u, float).reshape((Nx, Ny, 2, L, B), order="F"), np.float64)
    zz = _lib.f_array(np.asarray(z, float).reshape((2, 2, B), order="F"), np.float64)
    yy = _lib.f_array(np.asarray(y, float).reshape((2, L, B), order="F"), np.float64)
    du = np.empty_like(uu, order="F"); dz = np.empty_like(zz, order="F"); dy = np.empty_like(yy, order="F")
    ci, cm = _controls(qs, L)
    g = _lib.make_grid(p)
    pa = _params_array(qs, p)
    rc = ctx.lib.smm_pde_rhs(ctx.h, C.byref(g), pa, B, L, _lib.PREC[precision], _lib.ptr(ci), _lib.ptr(cm),
                             _lib.ptr(uu), _lib.ptr(zz), _lib.ptr(yy), _lib.ptr(du), _lib.ptr(dz), _lib.ptr(dy))
    ctx.check(rc, "smm_pde_rhs")
    if single:
        return du[..., 0], dz[..., 0], dy[..., 0]
    return du, dz, dy


def f(duzy, uzy, pq, t=0.0, *, precision="f32"):
    """The reference's in-place right-hand side f(duzy, uzy, (p,q), t) (pde.jl:183): fills duzy.x."""
    p, q = pq
    u, z, y = uzy.x
    du, dz, dy = pde_rhs(u, z, y, p=p, q=q, precision=precision)
    duzy.x[0][...] = du; duzy.x[1][...] = dz; duzy.x[2][...] = dy
    return None


class PDEBatch:
    """B independent PDE instances resident on the GPU (ensembles / parameter sweeps; C4)."""

    def __init__(self, p, qs, u0, z0, y0, *, precision="f32", shared_init=None, ctx=None):
        self.ctx = _lib.Context.default() if ctx is None else ctx
        self.p, self.qs = p, list(qs)
        B = self.B = len(self.qs)
        L = self.L = int(self.qs[0].L)
        Nx, Ny = p.N_x, p.N_y
        if shared_init is None:
            shared_init = np.ndim(u0) == 4 and B > 1
        nin = 1 if shared_init else B
        uu = _lib.f_array(np.asarray(u0, float).reshape((Nx, Ny, 2, L, nin), order="F"), np.float64)
        zz = _lib.f_array(np.asarray(z0, float).reshape((2, 2, nin), order="F"), np.float64)
        yy = _lib.f_array(np.asarray(y0, float).reshape((2, L, nin), order="F"), np.float64)
        ci, cm = _controls(self.qs, L)
        g = _lib.make_grid(p)
        pa = _params_array(self.qs, p)
        h = C.c_void_p()
        rc = self.ctx.lib.smm_pde_batch_create(self.ctx.h, C.byref(g), pa, B, L, _lib.PREC[precision], _lib.ptr(ci),
                                               _lib.ptr(cm), _lib.ptr(uu), _lib.ptr(zz), _lib.ptr(yy),
                                               1 if shared_init else 0, C.byref(h))
        self.ctx.check(rc, "smm_pde_batch_create")
        self.h = h

    def step(self, dt, nsteps=1, scheme="rk4"):
        self.ctx.check(self.ctx.lib.smm_pde_batch_step(self.ctx.h, self.h, float(dt), int(nsteps), _lib.SCHEME[scheme]),
                       "smm_pde_batch_step")

    def advance(self, t_end, *, abstol=1e-6, reltol=1e-3, dt0=0.0, dtmin=0.0, dtmax=0.0, force_dtmin=False,
                max_attempts=0):
        """Adaptive integration of every instance to the absolute time ``t_end`` (smm_pde_batch_advance; the
        tolerances / dtmin / force_dtmin keywords of the reference's solve calls, controlstrategies.jl:10,96).
        Returns (accepted, rejected, next_dt) per instance."""
        opt = _lib.make_adapt(abstol, reltol, dt0, dtmin, dtmax, force_dtmin, max_attempts)
        na = np.zeros(self.B, np.int64); nr = np.zeros(self.B, np.int64); dn = np.zeros(self.B)
        self.ctx.check(self.ctx.lib.smm_pde_batch_advance(self.ctx.h, self.h, float(t_end), C.byref(opt), _lib.ptr(na),
                                                          _lib.ptr(nr), _lib.ptr(dn)), "smm_pde_batch_advance")
        return na, nr, dn

    def read(self, want_u=True):
        Nx, Ny, L, B = self.p.N_x, self.p.N_y, self.L, self.B
        u = np.empty((Nx, Ny, 2, L, B), order="F") if want_u else None
        z = np.empty((2, 2, B), order="F"); y = np.empty((2, L, B), order="F")
        self.ctx.check(self.ctx.lib.smm_pde_batch_read(self.ctx.h, self.h, _lib.ptr(u), _lib.ptr(z), _lib.ptr(y)),
                       "smm_pde_batch_read")
        return u, z, y

    def sum_dev(self, sum_u_ptr, sum_z_ptr=None, sum_y_ptr=None):
        self.ctx.check(self.ctx.lib.smm_pde_batch_sum_dev(self.ctx.h, self.h, C.c_void_p(sum_u_ptr),
                                                          C.c_void_p(sum_z_ptr) if sum_z_ptr else None,
                                                          C.c_void_p(sum_y_ptr) if sum_y_ptr else None),
                       "smm_pde_batch_sum_dev")

    def close(self):
        if self.h:
            self.ctx.lib.smm_pde_batch_destroy(self.ctx.h, self.h)
            self.h = None


def default_dt(p, q):
    """Stable explicit step: diffusion D dt/dx^2 <= 0.16 and dt <= 1e-3 (SURVEY hard part 6)."""
    D = 0.5 * float(q.sigma) ** 2
    return min(1e-3, 0.16 * p.dx * p.dx / max(D, 1e-12))


class PDESolution:
    """What PDEsolve returns: ``sol.t`` (saved times), ``sol.u`` (list of ArrayPartition) and
    ``sol(t)`` = state at time t (linear interpolation between saved steps; exact on them)."""

    def __init__(self, t, states):
        self.t = np.asarray(t, float)
        self.u = states

    def __call__(self, t):
        t = float(t)
        k = int(np.searchsorted(self.t, t))
        if k < len(self.t) and abs(self.t[k] - t) <= 1e-12 * max(1.0, abs(t)):
            return self.u[k]
        if k == 0 or k >= len(self.t):
            raise ValueError(f"t={t} outside the solved interval [{self.t[0]}, {self.t[-1]}]")
        w = (t - self.t[k - 1]) / (self.t[k] - self.t[k - 1])
        a, b = self.u[k - 1].x, self.u[k].x
        return ArrayPartition(*[(1 - w) * ai + w * bi for ai, bi in zip(a, b)])


def pde_solve_raw(p, qs, u0, z0, y0, *, dt, nsteps, savesteps, scheme="rk4", precision="f32", want_sums=False, ctx=None):
    """smm_pde_solve: us[Nx,Ny,2,L,nsave,B], zs[2,2,nsave,B], ys[2,L,nsave,B] (+ sums over B)."""
    ctx = _lib.Context.default() if ctx is None else ctx
    qs = list(qs)
    B, L = len(qs), int(qs[0].L)
    Nx, Ny = p.N_x, p.N_y
    savesteps = np.ascontiguousarray(savesteps, np.int32)
    ns = len(savesteps)
    uu = _lib.f_array(np.asarray(u0, float).reshape((Nx, Ny, 2, L, B), order="F"), np.float64)
    zz = _lib.f_array(np.asarray(z0, float).reshape((2, 2, B), order="F"), np.float64)
    yy = _lib.f_array(np.asarray(y0, float).reshape((2, L, B), order="F"), np.float64)
    us = np.empty((Nx, Ny, 2, L, ns, B), order="F"); zs = np.empty((2, 2, ns, B), order="F"); ys = np.empty((2, L, ns, B), order="F")
    su = np.empty((Nx, Ny, 2, L, ns), order="F") if want_sums else None
    sz = np.empty((2, 2, ns), order="F") if want_sums else None
    sy = np.empty((2, L, ns), order="F") if want_sums else None
    ci, cm = _controls(qs, L)
    g = _lib.make_grid(p)
    pa = _params_array(qs, p)
    rc = ctx.lib.smm_pde_solve(ctx.h, C.byref(g), pa, B, L, _lib.PREC[precision], _lib.ptr(ci), _lib.ptr(cm),
                               _lib.ptr(uu), _lib.ptr(zz), _lib.ptr(yy), float(dt), int(nsteps), _lib.SCHEME[scheme],
                               ns, _lib.ptr(savesteps), _lib.ptr(us), _lib.ptr(zs), _lib.ptr(ys), _lib.ptr(su),
                               _lib.ptr(sz), _lib.ptr(sy))
    ctx.check(rc, "smm_pde_solve")
    out = dict(us=us, zs=zs, ys=ys)
    if want_sums:
        out.update(sum_us=su, sum_zs=sz, sum_ys=sy)
    return out


def pde_solve_adaptive_raw(p, qs, u0, z0, y0, *, savetimes, abstol=1e-6, reltol=1e-3, dt0=0.0, dtmin=0.0, dtmax=0.0,
                           force_dtmin=False, max_attempts=0, precision="f32", want_sums=False, ctx=None):
    """smm_pde_solve_adaptive: like pde_solve_raw, with states recorded at the times ``savetimes`` and
    per-instance accepted / rejected step counts in ``naccept`` / ``nreject``."""
    ctx = _lib.Context.default() if ctx is None else ctx
    qs = list(qs)
    B, L = len(qs), int(qs[0].L)
    Nx, Ny = p.N_x, p.N_y
    savetimes = np.ascontiguousarray(savetimes, np.float64)
    ns = len(savetimes)
    uu = _lib.f_array(np.asarray(u0, float).reshape((Nx, Ny, 2, L, B), order="F"), np.float64)
    zz = _lib.f_array(np.asarray(z0, float).reshape((2, 2, B), order="F"), np.float64)
    yy = _lib.f_array(np.asarray(y0, float).reshape((2, L, B), order="F"), np.float64)
    us = np.empty((Nx, Ny, 2, L, ns, B), order="F"); zs = np.empty((2, 2, ns, B), order="F"); ys = np.empty((2, L, ns, B), order="F")
    su = np.empty((Nx, Ny, 2, L, ns), order="F") if want_sums else None
    sz = np.empty((2, 2, ns), order="F") if want_sums else None
    sy = np.empty((2, L, ns), order="F") if want_sums else None
    na = np.zeros(B, np.int64); nr = np.zeros(B, np.int64)
    ci, cm = _controls(qs, L)
    g = _lib.make_grid(p)
    pa = _params_array(qs, p)
    opt = _lib.make_adapt(abstol, reltol, dt0, dtmin, dtmax, force_dtmin, max_attempts)
    rc = ctx.lib.smm_pde_solve_adaptive(ctx.h, C.byref(g), pa, B, L, _lib.PREC[precision], _lib.ptr(ci), _lib.ptr(cm),
                                        _lib.ptr(uu), _lib.ptr(zz), _lib.ptr(yy), C.byref(opt), ns, _lib.ptr(savetimes),
                                        _lib.ptr(us), _lib.ptr(zs), _lib.ptr(ys), _lib.ptr(su), _lib.ptr(sz), _lib.ptr(sy),
                                        _lib.ptr(na), _lib.ptr(nr))
    ctx.check(rc, "smm_pde_solve_adaptive")
    out = dict(us=us, zs=zs, ys=ys, naccept=na, nreject=nr, t=savetimes)
    if want_sums:
        out.update(sum_us=su, sum_zs=sz, sum_ys=sy)
    return out


def PDEsolve(tmax=0.1, *, alg=None, init="4inf", p=None, q=None, dt=None, scheme="rk4", precision="f32",
             nsaves=101, rng=None, ctx=None, adaptive=False, abstol=1e-6, reltol=1e-3, dtmin=0.0, force_dtmin=False,
             saveat=None):
    """pde.jl:285-294.  Returns (sol, (p, q)) with q augmented by controlled_inf / controlled_med as in the
    reference (pde.jl:178).  Integration: by default the fixed-step ``scheme`` with step ``dt`` (default:
    default_dt); with ``adaptive=True`` (or ``alg="BS3"``) the embedded 3(2) pair with per-instance step control
    and the reference's ``abstol`` / ``reltol`` / ``dtmin`` / ``force_dtmin`` / ``saveat`` keywords
    (controlstrategies.jl:10).  Any other ``alg`` is accepted for signature compatibility and ignored."""
    p = PDEconstruct() if p is None else p
    q = parameters() if q is None else q
    uzy0, q = constructinitial(init, (p, q), rng)
    if adaptive or (isinstance(alg, str) and alg.upper() in ("BS3", "ADAPTIVE")):
        times = np.linspace(0.0, tmax, nsaves) if saveat is None else np.asarray(saveat, float)
        u0, z0, y0 = uzy0.x
        out = pde_solve_adaptive_raw(p, [q], u0, z0, y0, savetimes=times, abstol=abstol, reltol=reltol, dtmin=dtmin,
                                     force_dtmin=force_dtmin, precision=precision, ctx=ctx)
        states = [ArrayPartition(out["us"][:, :, :, :, m, 0], out["zs"][:, :, m, 0], out["ys"][:, :, m, 0])
                  for m in range(len(times))]
        sol = PDESolution(times, states)
        sol.naccept, sol.nreject = int(out["naccept"][0]), int(out["nreject"][0])
        return sol, (p, q)
    dt = default_dt(p, q) if dt is None else dt
    nsteps = max(1, int(np.ceil(tmax / dt - 1e-9)))
    dt = tmax / nsteps
    savesteps = np.unique(np.round(np.linspace(0, nsteps, min(nsaves, nsteps + 1))).astype(np.int32))
    u0, z0, y0 = uzy0.x
    out = pde_solve_raw(p, [q], u0, z0, y0, dt=dt, nsteps=nsteps, savesteps=savesteps, scheme=scheme,
                        precision=precision, ctx=ctx)
    states = [ArrayPartition(out["us"][:, :, :, :, m, 0], out["zs"][:, :, m, 0], out["ys"][:, :, m, 0])
              for m in range(len(savesteps))]
    return PDESolution(savesteps * dt, states), (p, q)


def sol2uyz(sol, t):
    """pde.jl:277-282 (note the return order u, z, y)."""
    s = sol(t)
    return s.x[0], s.x[1], s.x[2]
