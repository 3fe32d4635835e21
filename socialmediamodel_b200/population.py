"""A large ABM population resident on the GPU (config C5: 1e7 agents, CSR follower network,
64 influencers) -- the scaled form of the loop body abm.jl:166-213, one fused kernel per step."""
import ctypes as C

import numpy as np

from . import _lib


class ABMPopulation:
    def __init__(self, q, p, x0, state, follow0, inf0, media0, network=None, *, seed=0, precision="f32", ctx=None):
        self.ctx = _lib.Context.default() if ctx is None else ctx
        self.q, self.p = q, p
        n, L, M = q.n, q.L, q.M
        x0 = _lib.f_array(np.asarray(x0, float).reshape((n, 2), order="F"), np.float64)
        st = np.ascontiguousarray(state, np.int32).reshape(n)
        f0 = np.ascontiguousarray(follow0, np.int32).reshape(n)
        i0 = _lib.f_array(np.asarray(inf0, float).reshape((L, 2), order="F"), np.float64)
        m0 = _lib.f_array(np.asarray(media0, float).reshape((M, 2), order="F"), np.float64)
        rowptr = colidx = None
        if network is not None:
            rowptr = np.ascontiguousarray(network[0], np.int32)
            colidx = np.ascontiguousarray(network[1], np.int32)
        sp = _lib.make_params(q, p)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.smm_abm_pop_create(self.ctx.h, C.byref(sp), _lib.PREC[precision], int(seed),
                                                       _lib.ptr(x0), _lib.ptr(st), _lib.ptr(f0), _lib.ptr(i0),
                                                       _lib.ptr(m0), _lib.ptr(rowptr), _lib.ptr(colidx), C.byref(h)),
                       "smm_abm_pop_create")
        self.h = h

    def step(self, nsteps=1):
        """Advance nsteps with Philox noise; asynchronous (call ctx.sync() or read())."""
        self.ctx.check(self.ctx.lib.smm_abm_pop_step(self.ctx.h, self.h, int(nsteps)), "smm_abm_pop_step")

    def step_injected(self, xi, u1, u2):
        n = self.q.n
        xi = _lib.f_array(np.asarray(xi, float).reshape((n, 2), order="F"), np.float64)
        u1 = np.ascontiguousarray(u1, np.float64); u2 = np.ascontiguousarray(u2, np.float64)
        self.ctx.check(self.ctx.lib.smm_abm_pop_step_injected(self.ctx.h, self.h, _lib.ptr(xi), _lib.ptr(u1),
                                                              _lib.ptr(u2)), "smm_abm_pop_step_injected")

    def read(self, positions=True):
        n, L, M = self.q.n, self.q.L, self.q.M
        x = np.empty((n, 2), order="F") if positions else None
        follow = np.empty(n, np.int32) if positions else None
        inf = np.empty((L, 2), order="F"); media = np.empty((M, 2), order="F")
        numbers = np.empty((M, L), np.int64, order="F")
        self.ctx.check(self.ctx.lib.smm_abm_pop_read(self.ctx.h, self.h, _lib.ptr(x), _lib.ptr(follow), _lib.ptr(inf),
                                                     _lib.ptr(media), _lib.ptr(numbers)), "smm_abm_pop_read")
        return dict(x=x, follow=follow, inf=inf, media=media, numbers=numbers)

    def stats(self):
        steps = C.c_int64(); fb = C.c_int64()
        self.ctx.check(self.ctx.lib.smm_abm_pop_stats(self.ctx.h, self.h, C.byref(steps), C.byref(fb)), "smm_abm_pop_stats")
        return dict(steps=steps.value, fp64_reevaluations=fb.value)

    def close(self):
        if self.h:
            self.ctx.lib.smm_abm_pop_destroy(self.ctx.h, self.h)
            self.h = None
