"""ABMsolve of the reference (src/abm.jl:154-217) on the B200 path.

Same positional/keyword signature and return value as the reference; extra keywords (with
defaults that reproduce the reference call) select the device arithmetic, inject the random
draws for parity runs, and pass a sparse follower network (the reference builds
``IndNet = ones(n,n)`` inside ABMinit and cannot take one, abm.jl:42).
"""
import ctypes as C

import numpy as np

from . import _lib
from .setting import ABMconstruct, parameters


def _noise_struct(seed, injected, n, L, M, NT, R):
    ns = _lib.smm_noise()
    keep = []
    ns.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if injected is None:
        ns.mode = _lib.NOISE_PHILOX
        return ns, keep
    ns.mode = _lib.NOISE_INJECTED
    xi = _lib.f_array(np.asarray(injected["xi"], float).reshape((n, 2, NT, R), order="F"), np.float64)
    u1 = _lib.f_array(np.asarray(injected["u1"], float).reshape((n, NT, R), order="F"), np.float64)
    u2 = _lib.f_array(np.asarray(injected["u2"], float).reshape((n, NT, R), order="F"), np.float64)
    keep += [xi, u1, u2]
    ns.xi, ns.u1, ns.u2 = xi.ctypes.data, u1.ctypes.data, u2.ctypes.data
    for key, cnt in (("xi_inf", L), ("xi_med", M)):
        if injected.get(key) is not None:
            a = _lib.f_array(np.asarray(injected[key], float).reshape((2, cnt, NT, R), order="F"), np.float64)
            keep.append(a)
            setattr(ns, key, a.ctypes.data)
    return ns, keep


def _network(network, n):
    if network is None:
        return None, None
    rowptr, colidx = network
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    assert rowptr.shape == (n + 1,)
    return rowptr, colidx


def abm_init(q=None, p=None, *, seed=0, first_realization=0, R=1, ctx=None):
    """ABMinit (abm.jl:8-45) from the device's Philox INIT stream.  Returns x0[n,2,R], state[n,R],
    follow0[n,R] (0-based), inf0[L,2,R], media0[M,2,R]."""
    q = parameters() if q is None else q
    p = ABMconstruct() if p is None else p
    ctx = _lib.Context.default() if ctx is None else ctx
    n, L, M = q.n, q.L, q.M
    x0 = np.empty((n, 2, R), order="F"); state = np.empty((n, R), np.int32, order="F")
    follow0 = np.empty((n, R), np.int32, order="F")
    inf0 = np.empty((L, 2, R), order="F"); media0 = np.empty((M, 2, R), order="F")
    sp = _lib.make_params(q, p)
    ctx.check(ctx.lib.smm_abm_init(ctx.h, C.byref(sp), int(seed), int(first_realization), int(R), _lib.ptr(x0),
                                   _lib.ptr(state), _lib.ptr(follow0), _lib.ptr(inf0), _lib.ptr(media0)), "smm_abm_init")
    return x0, state, follow0, inf0, media0


def abm_draws(n, NT, *, seed=0, realization=0, ctx=None):
    """The Philox draws of one realization: xi[n,2,NT], u1[n,NT], u2[n,NT]."""
    ctx = _lib.Context.default() if ctx is None else ctx
    xi = np.empty((n, 2, NT), order="F"); u1 = np.empty((n, NT), order="F"); u2 = np.empty((n, NT), order="F")
    ctx.check(ctx.lib.smm_abm_draws(ctx.h, int(n), int(seed), int(realization), int(NT), _lib.ptr(xi), _lib.ptr(u1),
                                    _lib.ptr(u2)), "smm_abm_draws")
    return xi, u1, u2


def abm_solve_raw(NT, *, p, q, R=1, first_realization=0, seed=0, precision="f32", init_state=None, injected=None,
                  network=None, savesteps=None, want=("xs", "follows", "infs", "meds", "numbers"), ctx=None):
    """Thin wrapper of smm_abm_solve returning Julia-shaped arrays:
    xs[n,2,nsave,R], follows[n,nsave,R] (0-based int32), infs[L,2,NT+1,R], meds[M,2,NT+1,R], numbers[M,L,NT+1]."""
    ctx = _lib.Context.default() if ctx is None else ctx
    n, L, M = q.n, q.L, q.M
    savesteps = np.arange(NT + 1, dtype=np.int32) if savesteps is None else np.ascontiguousarray(savesteps, np.int32)
    nsave = len(savesteps)
    out = {}
    if "xs" in want: out["xs"] = np.empty((n, 2, nsave, R), order="F")
    if "follows" in want: out["follows"] = np.empty((n, nsave, R), np.int32, order="F")
    if "infs" in want: out["infs"] = np.empty((L, 2, NT + 1, R), order="F")
    if "meds" in want: out["meds"] = np.empty((M, 2, NT + 1, R), order="F")
    if "numbers" in want: out["numbers"] = np.empty((M, L, NT + 1), np.int64, order="F")
    keep = []
    if init_state is not None:
        x0 = _lib.f_array(np.asarray(init_state["x0"], float).reshape((n, 2, R), order="F"), np.float64)
        st = _lib.f_array(np.asarray(init_state["state"]).reshape((n, R), order="F"), np.int32)
        f0 = _lib.f_array(np.asarray(init_state["follow0"]).reshape((n, R), order="F"), np.int32)
        i0 = _lib.f_array(np.asarray(init_state["inf0"], float).reshape((L, 2, R), order="F"), np.float64)
        m0 = _lib.f_array(np.asarray(init_state["media0"], float).reshape((M, 2, R), order="F"), np.float64)
        keep += [x0, st, f0, i0, m0]
    else:
        x0 = st = f0 = i0 = m0 = None
    rowptr, colidx = _network(network, n)
    ns, k2 = _noise_struct(seed, injected, n, L, M, NT, R)
    keep += k2
    sp = _lib.make_params(q, p)
    rc = ctx.lib.smm_abm_solve(ctx.h, C.byref(sp), _lib.PREC[precision], int(NT), int(R), int(first_realization),
                               _lib.ptr(x0), _lib.ptr(st), _lib.ptr(f0), _lib.ptr(i0), _lib.ptr(m0),
                               _lib.ptr(rowptr), _lib.ptr(colidx), C.byref(ns), nsave, _lib.ptr(savesteps),
                               _lib.ptr(out.get("xs")), _lib.ptr(out.get("follows")), _lib.ptr(out.get("infs")),
                               _lib.ptr(out.get("meds")), _lib.ptr(out.get("numbers")))
    ctx.check(rc, "smm_abm_solve")
    return out


def ABMsolve(NT=100, *, p=None, q=None, init="4inf", chosenseed=0, precision="f32", init_state=None, injected=None,
             network=None, ctx=None):
    """abm.jl:154-217.  Returns (xs, stateinfs, infs, meds, state, (p, q)) exactly shaped like the
    reference: lists of length NT+1 of x[n,2], Float64 follower ids 1.0..L (abm.jl:164,212), inf[L,2],
    media[M,2]; state[n] in {-1,+1}.  ``init`` is accepted and ignored, as in the reference."""
    p = ABMconstruct() if p is None else p
    q = parameters() if q is None else q
    ctx = _lib.Context.default() if ctx is None else ctx
    if init_state is None:
        x0, state, follow0, inf0, media0 = abm_init(q, p, seed=chosenseed, R=1, ctx=ctx)
        init_state = dict(x0=x0, state=state, follow0=follow0, inf0=inf0, media0=media0)
    out = abm_solve_raw(NT, p=p, q=q, R=1, seed=chosenseed, precision=precision, init_state=init_state,
                        injected=injected, network=network, ctx=ctx)
    # lists of per-step matrices as the reference returns them; the entries are column-major views of the history arrays
    # (Julia's layout), the follower ids are widened once for all steps
    X, I, Md = out["xs"][:, :, :, 0], out["infs"][:, :, :, 0], out["meds"][:, :, :, 0]
    F = np.asfortranarray(out["follows"][:, :, 0].astype(np.float64) + 1.0)
    xs = [X[:, :, k] for k in range(NT + 1)]
    stateinfs = [F[:, k] for k in range(NT + 1)]
    infs = [I[:, :, k] for k in range(NT + 1)]
    meds = [Md[:, :, k] for k in range(NT + 1)]
    state = np.asarray(init_state["state"]).reshape(-1).astype(np.int64)
    return xs, stateinfs, infs, meds, state, (p, q)
