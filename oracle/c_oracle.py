"""ctypes front end of oracle/abm_oracle.c (the plain-C restatement of the reference's ABM ensemble path).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Only tests/, __graft_entry__.smoke()/build() and bench.py's
CPU-baseline legs may use it.  The shared object is compiled on demand with gcc into oracle/_build/ (git-ignored;
__graft_entry__.build() compiles it so that it travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "abm_oracle.c")
OUT = os.path.join(HERE, "_build", "liboracle_c.so")


class Params(C.Structure):
    _fields_ = [("n", C.c_int), ("L", C.c_int), ("M", C.c_int)] + \
               [(k, C.c_double) for k in ("a", "b", "c", "eta", "sigma", "sigmahat", "sigmatilde", "frictionI", "frictionM", "dt")] + \
               [("lo", C.c_double * 2), ("hi", C.c_double * 2)]


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.smm_c_ensemble_bench.restype = C.c_double
    return _lib


def make_params(q, p):
    d = np.asarray(p["domain"], float)
    P = Params(int(q["n"]), int(q["L"]), int(q["M"]), *[float(q[k]) for k in ("a", "b", "c", "eta", "sigma", "sigmahat", "sigmatilde",
                                                                           "frictionI", "frictionM")], float(p["dt"]))
    P.lo[0], P.lo[1], P.hi[0], P.hi[1] = d[0, 0], d[1, 0], d[0, 1], d[1, 1]
    return P


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def abm_solve(NT, x0, state, follow0, inf0, media0, xi, u1, u2, p, q):
    """Same contract as abm_oracle.abm_solve (all-to-all network): dict xs, follows, infs, meds."""
    n, L, M = int(q["n"]), int(q["L"]), int(q["M"])
    P = make_params(q, p)
    c = lambda a, t: np.ascontiguousarray(a, t)
    xs = np.empty((NT + 1, n, 2)); fs = np.empty((NT + 1, n), np.int32)
    infs = np.empty((NT + 1, L, 2)); meds = np.empty((NT + 1, M, 2))
    args = [c(x0, np.float64), c(state, np.int32), c(follow0, np.int32), c(inf0, np.float64), c(media0, np.float64),
            c(xi, np.float64), c(u1, np.float64), c(u2, np.float64)]
    rc = lib().smm_c_abm_solve(C.byref(P), int(NT), *[_ptr(a) for a in args], _ptr(xs), _ptr(fs), _ptr(infs), _ptr(meds))
    assert rc == 0
    return dict(xs=xs, follows=fs.astype(np.int64), infs=infs, meds=meds)


def kde(x, follow, state, p, q, sigma=0.1):
    """us[N_x,N_y,2,L] as ensemble_oracle.kde_snapshot, by the reference's non-separated double loop."""
    P = make_params(q, p)
    X, Y = p["X"], p["Y"]
    Nx, Ny = X.shape
    us = np.empty((2, int(q["L"]), Nx, Ny))
    c = lambda a, t: np.ascontiguousarray(a, t)
    xx, ff, ss = c(x, np.float64), c(follow, np.int32), c(state, np.int32)
    lib().smm_c_kde(C.byref(P), _ptr(xx), _ptr(ff), _ptr(ss), Nx, Ny, C.c_double(X[0, 0]), C.c_double(Y[0, 0]),
                    C.c_double(p["dx"]), C.c_double(p["dy"]), C.c_double(sigma), _ptr(us))
    return np.transpose(us, (2, 3, 0, 1))


def ensemble_bench(NT, R, p, q, savesteps, first=1, sigma=0.1, threads=0):
    """Run R realizations of ABMsolveensemble on the host cores (OpenMP over realizations); returns
    (checksum, threads used)."""
    P = make_params(q, p)
    X, Y = p["X"], p["Y"]
    Nx, Ny = X.shape
    st = np.ascontiguousarray(savesteps, np.int32)
    used = C.c_int(0)
    tot = lib().smm_c_ensemble_bench(C.byref(P), int(NT), int(R), C.c_int64(first), len(st), _ptr(st), Nx, Ny,
                                     C.c_double(X[0, 0]), C.c_double(Y[0, 0]), C.c_double(p["dx"]), C.c_double(p["dy"]),
                                     C.c_double(sigma), int(threads), C.byref(used))
    return float(tot), int(used.value)
