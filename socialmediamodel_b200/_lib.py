"""ctypes binding of libsmm_b200.so (the C ABI in include/smm_b200.h).

This is the Python stand-in for the Julia ``ccall`` shim (julia/SocialMediaModelB200.jl): same
entry points, same memory layouts (column-major Float64 / Int32, i.e. NumPy arrays in Fortran
order with the Julia shapes).  There is NO CPU fallback: if the shared library is missing or no
CUDA device is visible, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMM_B200_LIB: another build of the same library (A/B timing of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("SMM_B200_LIB") or os.path.join(_HERE, "libsmm_b200.so")

PREC = {"f32": 0, "f64": 1, 0: 0, 1: 1}
SCHEME = {"euler": 0, "heun": 1, "ssprk3": 2, "rk4": 3}
NOISE_PHILOX, NOISE_INJECTED = 0, 1


class SmmError(RuntimeError):
    pass


class smm_params(C.Structure):
    _fields_ = [("n", C.c_int32), ("L", C.c_int32), ("M", C.c_int32),
                ("eta", C.c_double), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("sigma", C.c_double), ("sigmahat", C.c_double), ("sigmatilde", C.c_double),
                ("frictionI", C.c_double), ("frictionM", C.c_double), ("dt", C.c_double),
                ("dom_lo", C.c_double * 2), ("dom_hi", C.c_double * 2),
                ("controlspeed1", C.c_double), ("controltarget1", C.c_double * 2),
                ("controlspeed2", C.c_double), ("controltarget2", C.c_double * 2)]


class smm_noise(C.Structure):
    _fields_ = [("mode", C.c_int32), ("seed", C.c_uint64), ("xi", C.c_void_p), ("u1", C.c_void_p),
                ("u2", C.c_void_p), ("xi_inf", C.c_void_p), ("xi_med", C.c_void_p)]


class smm_grid(C.Structure):
    _fields_ = [("Nx", C.c_int32), ("Ny", C.c_int32), ("x0", C.c_double), ("y0", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double)]


class smm_adapt(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("dt0", C.c_double), ("dtmin", C.c_double),
                ("dtmax", C.c_double), ("force_dtmin", C.c_int32), ("max_attempts", C.c_int32)]


def make_adapt(abstol=1e-6, reltol=1e-3, dt0=0.0, dtmin=0.0, dtmax=0.0, force_dtmin=False, max_attempts=0):
    """Step-control options of the adaptive integrator; the defaults are DifferentialEquations.jl's."""
    return smm_adapt(float(abstol), float(reltol), float(dt0), float(dtmin), float(dtmax), int(bool(force_dtmin)),
                     int(max_attempts))


_P = C.c_void_p
_I32, _I64, _U64, _D = C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes); every symbol declared in include/smm_b200.h
PROTOTYPES = {
    "smm_version": (C.c_int, []),
    "smm_device_count": (C.c_int, []),
    "smm_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "smm_ctx_create_multi": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "smm_ctx_gpu_count": (C.c_int, [_P]),
    "smm_ctx_nccl_active": (C.c_int, [_P]),
    "smm_ctx_destroy": (C.c_int, [_P]),
    "smm_ctx_sync": (C.c_int, [_P]),
    "smm_ctx_set_stream": (C.c_int, [_P, _P]),
    "smm_last_error": (C.c_char_p, [_P]),
    "smm_host_alloc": (C.c_int, [_P, _U64, C.POINTER(_P)]),
    "smm_host_free": (C.c_int, [_P, _P]),
    "smm_ctx_launch_count": (_I64, [_P]),
    "smm_ctx_kernel_ms": (_D, [_P, C.c_int]),
    "smm_microbench": (C.c_int, [_P, _I32, _U64, C.POINTER(_D)]),
    "smm_abm_init": (C.c_int, [_P, C.POINTER(smm_params), _U64, _I64, _I32, _P, _P, _P, _P, _P]),
    "smm_abm_draws": (C.c_int, [_P, _I32, _U64, _I64, _I32, _P, _P, _P]),
    "smm_abm_solve": (C.c_int, [_P, C.POINTER(smm_params), _I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P,
                                C.POINTER(smm_noise), _I32, _P, _P, _P, _P, _P, _P]),
    "smm_abm_ensemble": (C.c_int, [_P, C.POINTER(smm_params), _I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P,
                                   C.POINTER(smm_noise), _I32, _P, _D, C.POINTER(smm_grid),
                                   _P, _P, _P, _P, _P, _P, _P]),
    "smm_abm_ensemble_f32": (C.c_int, [_P, C.POINTER(smm_params), _I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P,
                                       C.POINTER(smm_noise), _I32, _P, _D, C.POINTER(smm_grid),
                                       _P, _P, _P, _P, _P, _P, _P]),
    "smm_abm_ensemble_dev": (C.c_int, [_P, C.POINTER(smm_params), _I32, _I32, _I32, _I64, _U64, _P, _P, _I32, _P, _D,
                                       C.POINTER(smm_grid), _P, _P, _P, _P, _P]),
    "smm_kde": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P, _D, C.POINTER(smm_grid), _P]),
    "smm_abm_pop_create": (C.c_int, [_P, C.POINTER(smm_params), _I32, _U64, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "smm_abm_pop_step": (C.c_int, [_P, _P, _I32]),
    "smm_abm_pop_step_injected": (C.c_int, [_P, _P, _P, _P, _P]),
    "smm_abm_pop_read": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "smm_abm_pop_stats": (C.c_int, [_P, _P, C.POINTER(_I64), C.POINTER(_I64)]),
    "smm_abm_pop_destroy": (C.c_int, [_P, _P]),
    "smm_pde_rhs": (C.c_int, [_P, C.POINTER(smm_grid), C.POINTER(smm_params), _I32, _I32, _I32, _P, _P,
                              _P, _P, _P, _P, _P, _P]),
    "smm_pde_solve": (C.c_int, [_P, C.POINTER(smm_grid), C.POINTER(smm_params), _I32, _I32, _I32, _P, _P,
                                _P, _P, _P, _D, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "smm_pde_batch_create": (C.c_int, [_P, C.POINTER(smm_grid), C.POINTER(smm_params), _I32, _I32, _I32, _P, _P,
                                       _P, _P, _P, _I32, C.POINTER(_P)]),
    "smm_pde_batch_step": (C.c_int, [_P, _P, _D, _I32, _I32]),
    "smm_pde_batch_read": (C.c_int, [_P, _P, _P, _P, _P]),
    "smm_pde_batch_sum_dev": (C.c_int, [_P, _P, _P, _P, _P]),
    "smm_pde_batch_destroy": (C.c_int, [_P, _P]),
    "smm_pde_batch_advance": (C.c_int, [_P, _P, _D, C.POINTER(smm_adapt), _P, _P, _P]),
    "smm_pde_solve_adaptive": (C.c_int, [_P, C.POINTER(smm_grid), C.POINTER(smm_params), _I32, _I32, _I32, _P, _P,
                                         _P, _P, _P, C.POINTER(smm_adapt), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load libsmm_b200.so (once).  Raises SmmError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmmError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C socialmediamodel_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    """Pointer to a NumPy array's data, or NULL."""
    return None if a is None else C.c_void_p(a.ctypes.data)


def f_array(a, dtype, shape=None):
    """Column-major (Julia layout) contiguous copy/view of `a`."""
    out = np.asfortranarray(a, dtype=dtype)
    if shape is not None:
        assert out.shape == tuple(shape), f"expected shape {shape}, got {out.shape}"
    return out


def make_params(q, p=None):
    """smm_params from the reference's q (setting.jl:20) and the ABM/PDE p (setting.jl:41,64)."""
    s = smm_params()
    s.n, s.L, s.M = int(q.n), int(q.L), int(q.M)
    for k in ("eta", "a", "b", "c", "sigma", "sigmahat", "sigmatilde", "frictionI", "frictionM",
              "controlspeed1", "controlspeed2"):
        setattr(s, k, float(getattr(q, k)))
    t1 = np.ravel(np.asarray(q.controltarget1, float)); t2 = np.ravel(np.asarray(q.controltarget2, float))
    s.controltarget1[0], s.controltarget1[1] = t1[0], t1[1]
    s.controltarget2[0], s.controltarget2[1] = t2[0], t2[1]
    dom = np.asarray(p.domain, float) if p is not None else 2.2 * np.array([[-1.0, 1.0], [-1.0, 1.0]])
    s.dt = float(getattr(p, "dt", 0.01)) if p is not None else 0.01
    s.dom_lo[0], s.dom_lo[1] = dom[0, 0], dom[1, 0]
    s.dom_hi[0], s.dom_hi[1] = dom[0, 1], dom[1, 1]
    return s


def make_grid(p):
    g = smm_grid()
    g.Nx, g.Ny = int(p.X.shape[0]), int(p.X.shape[1])
    g.x0, g.y0 = float(p.domain[0][0]), float(p.domain[1][0])
    g.dx, g.dy = float(p.dx), float(p.dy)
    return g


class Context:
    """One smm_ctx = one GPU.  ``Context.default()`` returns the per-process context on
    cuda:LOCAL_RANK (one process per GPU)."""
    _default = None

    def __init__(self, device=0, devices=None):
        """``devices``: None = one GPU (``device``); "all" or a list of device ids = several GPUs behind this one
        context (smm_ctx_create_multi): the ensemble calls shard over them inside the call."""
        self.lib = load()
        h = _P()
        if devices is None:
            rc = self.lib.smm_ctx_create(int(device), C.byref(h))
            what = f"smm_ctx_create({device})"
        else:
            ids = None if isinstance(devices, str) else np.ascontiguousarray(devices, np.int32)
            rc = self.lib.smm_ctx_create_multi(0 if ids is None else len(ids), ptr(ids), C.byref(h))
            what = f"smm_ctx_create_multi({devices})"
            device = 0 if ids is None else int(ids[0])
        if rc != 0:
            raise SmmError(f"{what} failed [{rc}]: {self.lib.smm_last_error(None).decode()}")
        self.h = h
        self.device = device
        self._pinned = {}

    @property
    def gpu_count(self):
        return int(self.lib.smm_ctx_gpu_count(self.h))

    @property
    def nccl_active(self):
        return bool(self.lib.smm_ctx_nccl_active(self.h))

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls(int(os.environ.get("LOCAL_RANK", "0")))
        return cls._default

    def check(self, rc, what):
        if rc != 0:
            raise SmmError(f"{what} failed [{rc}]: {self.lib.smm_last_error(self.h).decode()}")

    def sync(self):
        self.check(self.lib.smm_ctx_sync(self.h), "smm_ctx_sync")

    def set_stream(self, cuda_stream):
        self.check(self.lib.smm_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream))), "smm_ctx_set_stream")

    @property
    def launches(self):
        return int(self.lib.smm_ctx_launch_count(self.h))

    def kernel_ms(self, which):
        return float(self.lib.smm_ctx_kernel_ms(self.h, which))

    def microbench(self, which, size=0):
        """smm_microbench: measured MUFU / gather / shared-memory / PCIe / host-widening rates (see include/smm_b200.h)."""
        ids = dict(mufu=0, gather=1, lds=2, d2h=3, widen=4)
        r = _D(0.0)
        self.check(self.lib.smm_microbench(self.h, ids[which], int(size), C.byref(r)), f"smm_microbench({which})")
        return float(r.value)

    def pinned_empty(self, shape, dtype=np.float64, order="F"):
        """NumPy array backed by pinned host memory from smm_host_alloc (DMA target)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = _P()
        self.check(self.lib.smm_host_alloc(self.h, nbytes, C.byref(p)), "smm_host_alloc")
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape, order=order)
        self._pinned[arr.ctypes.data] = p
        return arr

    def free_pinned(self, arr):
        p = self._pinned.pop(arr.ctypes.data)
        self.check(self.lib.smm_host_free(self.h, p), "smm_host_free")

    def close(self):
        if self.h:
            self.lib.smm_ctx_destroy(self.h)
            self.h = None
