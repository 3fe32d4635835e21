"""The control layer as a caller of the GPU right-hand side (controlstrategies.jl:5-132):
parity of prep / PDEsolvefixedtargetsfast / PDEsolvefixedtargets against the oracle restatement running the
same fixed-step RK4, for the three countercontrol modes, and batched evaluation of candidate strategies."""
import numpy as np
import pytest

import socialmediamodel_b200 as smm
from oracle import abm_oracle as A
from oracle import control_oracle as CO
from oracle import pde_oracle as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cc,start", [("no", "zero"), ("inf", "influencer"), ("med", "mean")])
def test_fixed_targets_parity(cc, start):
    p = smm.PDEconstructcoarse()
    po = P.pde_construct_coarse()
    q = smm.parametersstronginf()
    qo = A.default_q(eta=1.0)
    r = smm.parameterscontrol(ntarg=2, start=start, Tmax=0.04, tequil=0.02)
    targets = np.array([[0.3, 0.3], [-0.5, 0.8]]).T                      # (2, ntarg)
    fs, sp, sols, Ps = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), p=p, q=q, r=r, countercontrol=cc,
                                                stubborntarget=[1.5, 1.5], savedt=0.01, dt=1e-3, precision="f64")
    uzy0, sl, q2, ci2, cm2, slm = CO.prep(0.02, po, qo, start=start, savedt=0.01, dt=1e-3, countercontrol=cc)
    ts = np.array([0.0, 0.02, 0.04])
    fr, pr, (ur, zr, yr) = CO.fixedtargets(ts, targets, sl, po, q2, ci2, cm2, uzy0, savedt=0.01, dt=1e-3, countercontrol=cc,
                                           stubborntarget=np.array([1.5, 1.5]), startlocationmed=slm)
    assert abs(fs - fr) < 1e-10 * max(1.0, abs(fr)) + 1e-12
    assert abs(sp - pr) < 1e-12
    uf, zf, yf = sols[-1](0.02).x
    assert np.abs(uf - ur).max() / np.abs(ur).max() < 1e-10
    np.testing.assert_allclose(yf, yr, atol=1e-11)
    np.testing.assert_allclose(zf, zr, atol=1e-11)
    assert len(sols) == 3 and Ps[1][1].L == (6 if cc == "inf" else 5)
    # the controlled entities move to their targets (pde.jl:243-261) -- up to one RK4 step of travel: the last
    # stage of the arriving step lands exactly on the target, where the reference defines the velocity as 0
    step_travel = np.linalg.norm(targets[:, 1] - targets[:, 0]) / 0.02 * 1e-3
    if cc == "inf":
        assert np.linalg.norm(yf[:, 4] - targets[:, 1]) < step_travel             # flag 1 -> the last target
    if cc == "med":
        assert np.linalg.norm(zf[:, 1] - targets[:, 1]) < step_travel             # controlled medium
    # FP32 run of the same evaluation stays within the north-star tolerance
    fs32, sp32, sols32, _ = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), p=p, q=q, r=r, countercontrol=cc,
                                                     stubborntarget=[1.5, 1.5], savedt=0.01, dt=1e-3, precision="f32")
    assert abs(fs32 - fr) < 1e-5 * max(abs(fr), 1e-3) + 1e-9
    assert np.abs(sols32[-1](0.02).x[0] - ur).max() / np.abs(ur).max() < 1e-5


def test_batched_candidate_strategies_equal_single_evaluations():
    """K candidate target sets evaluated as one device batch == K separate evaluations."""
    p = smm.PDEconstructcoarse()
    q = smm.parametersstronginf()
    r = smm.parameterscontrol(ntarg=2, start="zero", Tmax=0.04, tequil=0.02)
    uzy0, sl, (p, q2), slm = smm.prep(0.02, p=p, q=q, r=r, savedt=0.01, dt=1e-3, countercontrol="inf", precision="f64")
    ts = [0.0, 0.02, 0.04]
    cands = [np.array([[0.3, 0.3], [-0.5, 0.8]]).T, np.array([[1.0, -1.0], [0.0, 0.0]]).T, np.array([[-1.5, 1.5], [1.5, 1.5]]).T]
    fb, pb = smm.fixedtargets_objectives(ts, cands, sl, (p, q2), r, uzy0, savedt=0.01, dt=1e-3, countercontrol="inf",
                                         stubborntarget=[1.5, 1.5], precision="f64")
    for k, c in enumerate(cands):
        f1, p1 = smm.PDEsolvefixedtargetsfast(ts, c, sl, (p, q2), r, uzy0, savedt=0.01, dt=1e-3, countercontrol="inf",
                                              stubborntarget=[1.5, 1.5], precision="f64")
        assert abs(fb[k] - f1) < 1e-13 and abs(pb[k] - p1) < 1e-13
    assert len(set(np.round(fb, 12))) == 3                                 # different strategies, different objectives


@pytest.mark.parametrize("cc", ["no", "inf", "med"])
def test_adaptive_integrator_with_reference_keywords(cc):
    """integrator="adaptive": the reference's atol / rtol / dtmin / force_dtmin semantics
    (controlstrategies.jl:10,96) on the device-side embedded pair.  The objective agrees with the fixed-step
    evaluation within the requested tolerance, and a batch of candidates with its single evaluations."""
    p = smm.PDEconstructcoarse()
    q = smm.parametersstronginf()
    r = smm.parameterscontrol(ntarg=2, start="zero", Tmax=0.1, tequil=0.05, dtmin=1e-4)
    targets = np.array([[0.3, 0.3], [-0.5, 0.8]]).T
    kw = dict(p=p, q=q, r=r, countercontrol=cc, stubborntarget=[1.5, 1.5], savedt=0.01)
    ff, sf, solf, _ = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), dt=2.5e-4, precision="f64", **kw)
    for precision in ("f64", "f32"):
        fa, sa, sola, _ = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), precision=precision, integrator="adaptive", **kw)
        assert abs(sa - sf) < 1e-4          # "med": the start location is the equilibrated medium position
        assert abs(fa - ff) < 5e-3 * abs(ff) + 1e-6, (fa, ff)
        ua, uf = sola[-1](0.05).x[0], solf[-1](0.05).x[0]
        assert np.abs(ua - uf).max() / np.abs(uf).max() < 1e-2
        assert np.abs(sola[-1](0.05).x[2] - solf[-1](0.05).x[2]).max() < 1e-2
    uzy0, sl, (p2, q2), slm = smm.prep(0.05, p=p, q=q, r=r, savedt=0.01, countercontrol=cc, precision="f64", integrator="adaptive",
                                       dtmin=1e-4)
    ts = [0.0, 0.05, 0.1]
    cands = [targets, np.array([[1.0, -1.0], [0.0, 0.0]]).T]
    fb, pb = smm.fixedtargets_objectives(ts, cands, sl, (p2, q2), r, uzy0, savedt=0.01, countercontrol=cc, dtmin=1e-4,
                                         stubborntarget=[1.5, 1.5], startlocationmed=slm, precision="f64", integrator="adaptive")
    for k, c in enumerate(cands):
        f1, p1 = smm.PDEsolvefixedtargetsfast(ts, c, sl, (p2, q2), r, uzy0, savedt=0.01, countercontrol=cc, dtmin=1e-4,
                                              stubborntarget=[1.5, 1.5], startlocationmed=slm, precision="f64", integrator="adaptive")
        assert abs(fb[k] - f1) < 1e-9 and abs(pb[k] - p1) < 1e-13


def test_PDEsolveopt_batched_compass_search(tmp_path, monkeypatch):
    """PDEsolveopt (controlstrategies.jl:137-209) on the batched objective: improves on the starting point, stays
    inside the domain bounds, honours meval, records every evaluation, keeps the objective's alpha quirk."""
    monkeypatch.chdir(tmp_path)
    p = smm.PDEconstructcoarse()
    q = smm.parametersstronginf()
    r = smm.parameterscontrol(ntarg=1, start="zero", Tmax=0.1, tequil=0.02, alpha=0.05)
    x0 = np.array([0.2, -0.1])
    maxf, maxx, ret, numevals, xs, fsl, penl = smm.PDEsolveopt(p=p, q=q, r=r, x0=x0, countercontrol="inf", stubborntarget=[1.5, 1.5],
                                                               meval=30, savedt=0.02, dt=1e-3, precision="f64", save=True)
    assert numevals == len(xs) == len(fsl) == len(penl) <= 30 and numevals >= 5
    obj = np.array(fsl) - r.alpha * np.array(penl)                       # follower: alpha enters twice (:156,165)
    assert abs(maxf - obj.max()) < 1e-15 and maxf > obj[0]
    assert np.all(np.abs(maxx) <= p.domain[0, 1]) and np.all(np.abs(np.array(xs)) <= p.domain[0, 1] + 1e-12)
    np.testing.assert_array_equal(xs[0], x0)
    # the recorded objective of the best point equals a fresh single evaluation
    uzy0, sl, (p2, q2), slm = smm.prep(0.02, p=p, q=q, r=r, savedt=0.02, dt=1e-3, countercontrol="inf", precision="f64")
    f1, s1 = smm.PDEsolvefixedtargetsfast([0.0, 0.1], maxx.reshape(2, 1), sl, (p2, q2), r, uzy0, savedt=0.02, dt=1e-3,
                                          countercontrol="inf", stubborntarget=[1.5, 1.5], precision="f64")
    assert abs((f1 - r.alpha * r.alpha * s1) - maxf) < 1e-12
    assert ret in ("MAXEVAL_REACHED", "XTOL_REACHED", "FTOL_REACHED")
    import os
    assert os.path.exists("data/opt_strategy1LN_COBYLAfollowerzerofalseinf.npz")


def test_segments_that_are_not_multiples_of_savedt():
    """The reference's defaults (Tmax / ntarg = 5/3, savedt = 0.05) give control windows that are not multiples of the
    save interval: the objective sums over the save points k*savedt <= Dt (saveat = 0:savedt:Dt) and the next window
    starts from the state at Dt.  Float64 parity with the oracle, both integrators consistent."""
    p = smm.PDEconstructcoarse()
    po = P.pde_construct_coarse()
    q = smm.parametersstronginf()
    qo = A.default_q(eta=1.0)
    r = smm.parameterscontrol(ntarg=2, start="zero", Tmax=0.05, tequil=0.025)
    targets = np.array([[0.3, 0.3], [-0.5, 0.8]]).T
    fs, sp, sols, Ps = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), p=p, q=q, r=r, countercontrol="inf",
                                                stubborntarget=[1.5, 1.5], savedt=0.01, dt=1e-3, precision="f64")
    uzy0, sl, q2, ci2, cm2, slm = CO.prep(0.025, po, qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="inf")
    fr, pr, (ur, zr, yr) = CO.fixedtargets(np.array([0.0, 0.025, 0.05]), targets, sl, po, q2, ci2, cm2, uzy0, savedt=0.01,
                                           dt=1e-3, countercontrol="inf", stubborntarget=np.array([1.5, 1.5]))
    assert abs(fs - fr) < 1e-10 * max(1.0, abs(fr)) + 1e-12 and abs(sp - pr) < 1e-12
    assert [len(s.t) for s in sols] == [3, 3, 3]                     # 0, 0.01, 0.02 in every window of 0.025
    fa, sa, _, _ = smm.PDEsolvefixedtargets(targets.reshape(-1, order="F"), p=p, q=q, r=r, countercontrol="inf",
                                            stubborntarget=[1.5, 1.5], savedt=0.01, precision="f64", integrator="adaptive")
    assert abs(fa - fs) < 5e-3 * abs(fs) + 1e-7 and abs(sa - sp) < 1e-6
