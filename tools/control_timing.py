#!/usr/bin/env python
"""Wall time of one full-size control evaluation (controlstrategies.jl:120-132 with the reference's defaults:
coarse grid, strong influencers, tequil = 5, Tmax = 5, 3 targets) on the fixed-step and the adaptive integrator,
and of a batch of 16 candidate strategies evaluated at once.  GPU only."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import socialmediamodel_b200 as smm  # noqa: E402


def main():
    p, q, r = smm.PDEconstructcoarse(), smm.parametersstronginf(), smm.parameterscontrol()
    targets = np.array([0.5, 0.5, 1.0, 1.0, 1.5, 1.5])
    out = {}
    for integ in ("fixed", "adaptive"):
        smm.PDEsolvefixedtargets(targets, p=p, q=q, r=smm.parameterscontrol(Tmax=0.3, tequil=0.1), countercontrol="inf", integrator=integ)
        t0 = time.perf_counter()
        fs, pen, sols, Ps = smm.PDEsolvefixedtargets(targets, p=p, q=q, r=r, countercontrol="inf", integrator=integ)
        out[integ] = dict(seconds=time.perf_counter() - t0, followersum=fs, speedpenalty=pen,
                          mass=float(sols[-1].u[-1].x[0].sum() * p.dV))
    uzy0, sl, (p2, q2), slm = smm.prep(r.tequil, p=p, q=q, r=r, countercontrol="inf", integrator="adaptive")
    ts = np.concatenate([[0.0], r.Tmax / r.ntarg * np.arange(1, r.ntarg + 1)])
    rng = np.random.default_rng(0)
    cands = [rng.uniform(-2, 2, (2, r.ntarg)) for _ in range(16)]
    t0 = time.perf_counter()
    fb, pb = smm.fixedtargets_objectives(ts, cands, sl, (p2, q2), r, uzy0, countercontrol="inf", stubborntarget=[1.5, 1.5],
                                         integrator="adaptive")
    out["batch16_adaptive"] = dict(seconds=time.perf_counter() - t0, best=float(fb.max()), worst=float(fb.min()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
