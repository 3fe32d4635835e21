"""The oracle must keep reproducing the committed golden fixtures (tests/golden/*.npz, written by
tests/golden/make_golden.py): guards the restatement of the reference against silent drift."""
import os

import numpy as np

from oracle import abm_oracle as A
from oracle import ensemble_oracle as E
from oracle import pde_oracle as P

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_abm_goldens_reproduced_by_oracle():
    for name in ("abm_dense_n60.npz", "abm_csr_n80.npz"):
        g = np.load(os.path.join(G, name))
        n, NT = int(g["n"]), int(g["NT"])
        net = (g["rowptr"], g["colidx"]) if "rowptr" in g else None
        o = A.abm_solve(NT, g["x0"], g["state"], g["follow0"], g["inf0"], g["media0"], g["xi"], g["u1"], g["u2"],
                        A.default_p(), A.default_q(n=n), net)
        np.testing.assert_array_equal(o["follows"], g["follows"])
        np.testing.assert_allclose(o["xs"], g["xs"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(o["infs"], g["infs"], rtol=1e-13)
        p = A.default_p()
        t = g["savesteps"][-1]
        us = E.kde_snapshot(o["xs"][t], o["follows"][t], g["state"], 4, p["X"], p["Y"], n)
        np.testing.assert_allclose(us, g["us"][..., -1], rtol=1e-12, atol=1e-15)
        assert (np.diff(g["follows"], axis=0) != 0).sum() > 5          # the fixture exercises switching


def test_pde_goldens_reproduced_by_oracle():
    g = np.load(os.path.join(G, "pde_coarse.npz"))
    p, q = P.pde_construct_coarse(), A.default_q()
    du, dz, dy = P.rhs(g["u"], g["z"], g["y"], p, q)
    np.testing.assert_allclose(du, g["du"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(dy, g["dy"], rtol=1e-11, atol=1e-14)
    du5, dz5, dy5 = P.rhs(g["u5"], g["z"], g["y5"], p, dict(q, L=5), g["ci5"], g["cm5"])
    np.testing.assert_allclose(du5, g["du5"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(dz5, g["dz5"], rtol=1e-12)


def test_adaptive_and_control_goldens_reproduced_by_oracle():
    from oracle import control_oracle as CO
    g = np.load(os.path.join(G, "pde_adaptive_control.npz"))
    p, q = P.pde_construct_coarse(), A.default_q()
    u0, z0, y0, ci, cm = P.uniforminit(p, q)
    states, log = P.bs3_adaptive(u0, z0, y0, g["save"], p, q)
    got = np.array([(t, dt, n, float(a)) for t, dt, n, a in log])
    assert got.shape == g["log"].shape
    np.testing.assert_array_equal(got[:, 3], g["log"][:, 3])                 # accept / reject sequence
    np.testing.assert_allclose(got[:, :2], g["log"][:, :2], rtol=1e-12)      # times and step sizes
    np.testing.assert_allclose(np.stack([s[0] for s in states], axis=-1), g["u_adapt"], rtol=1e-11, atol=1e-14)
    qo = A.default_q(eta=1.0)
    uzy0, sl, q2, ci2, cm2, slm = CO.prep(0.02, p, qo, start="zero", savedt=0.01, dt=1e-3, countercontrol="inf")
    fs, pen, (uf, zf, yf) = CO.fixedtargets(g["ctrl_ts"], g["ctrl_targets"], sl, p, q2, ci2, cm2, uzy0, savedt=0.01, dt=1e-3,
                                            countercontrol="inf", stubborntarget=np.array([1.5, 1.5]))
    assert abs(fs - float(g["ctrl_followersum"])) < 1e-13 and abs(pen - float(g["ctrl_penalty"])) < 1e-13
    np.testing.assert_allclose(yf, g["ctrl_y_final"], rtol=1e-12)


# ---- the reference itself (fixtures written by baseline/julia_fixtures.jl; absent in this image: no Julia) -------------
import pytest

import ref_fixtures as RF

needs_ref = pytest.mark.skipif(not RF.available(), reason=RF.SKIP_REASON)


def test_reference_log_replay_roundtrip():
    """The replay of the reference's draw logs (ref_fixtures.replay_abm) is exercised without Julia: logs of the layout
    baseline/julia_fixtures.jl writes are emulated from an oracle run and must replay to the same run, consuming every
    draw exactly once (a second uniform only for agents that enter the switch branch, abm.jl:122-126)."""
    p, q = A.default_p(), A.default_q(n=40)
    uni, nrm, xs, fs = RF.emulate_reference_log(3, 25, p, q)
    lu, ln = RF.Log(uni), RF.Log(nrm)
    r = RF.replay_abm(lu, ln, 25, p, q)
    assert lu.i == uni.size and ln.i == nrm.size
    np.testing.assert_array_equal(r["follows"], fs)
    np.testing.assert_array_equal(r["xs"], xs)
    assert (np.diff(fs, axis=0) != 0).sum() > 3 and uni.size > 3 * 40 + 25 * 40          # some agents switched


@needs_ref
@pytest.mark.parametrize("tag", ["abm_a", "abm_b", "abm_c"])
def test_oracle_against_reference_abmsolve(tag):
    """ABMsolve of the reference (abm.jl:154-217), its draws logged: follower histories identical, positions /
    influencers / media to rounding."""
    m = RF.manifest()[tag]
    p, q = A.default_p(), A.default_q()
    assert m["n"] == q["n"] and m["L"] == q["L"] and abs(m["dt"] - p["dt"]) < 1e-15
    lu, ln = RF.Log(RF.load(tag + "_uniform_log")), RF.Log(RF.load(tag + "_normal_log"))
    r = RF.replay_abm(lu, ln, m["NT"], p, q)
    assert lu.i == lu.v.size and ln.i == ln.v.size, "the oracle consumed a different number of draws than the reference"
    np.testing.assert_array_equal(r["state"], RF.load(tag + "_state"))
    np.testing.assert_array_equal(r["follows"].T + 1, RF.load(tag + "_stateinfs"))        # abm.jl:164,212: 1.0 .. L
    np.testing.assert_allclose(np.transpose(r["xs"], (1, 2, 0)), RF.load(tag + "_xs"), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.transpose(r["infs"], (1, 2, 0)), RF.load(tag + "_infs"), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.transpose(r["meds"], (1, 2, 0)), RF.load(tag + "_meds"), rtol=1e-12, atol=1e-13)


@needs_ref
def test_oracle_against_reference_forces():
    x, state, inf, media = RF.load("force_x"), RF.load("force_state").astype(int), RF.load("force_inf"), RF.load("force_media")
    follow = np.argmax(RF.load("force_FolInfNet"), axis=1)
    q = A.default_q()
    np.testing.assert_allclose(A.attraction(x, None), RF.load("force_attraction"), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(A.influence(x, media, inf, follow, state, q["b"], q["c"]), RF.load("force_influence"), rtol=1e-13, atol=1e-14)


@needs_ref
def test_oracle_against_reference_ensemble_kde():
    m = RF.manifest()["ens"]
    p = A.default_p(dx=m["dx"])
    q = A.default_q()
    lu, ln = RF.Log(RF.load("ens_uniform_log")), RF.Log(RF.load("ens_normal_log"))
    us, ys = RF.load("ens_us"), RF.load("ens_ys")
    st = E.savetimes_abm(m["NT"], m["savepoints"])
    for k in range(m["N"]):
        r = RF.replay_abm(lu, ln, m["NT"], p, q)
        for mi, t in enumerate(st):
            got = E.kde_snapshot(r["xs"][t], r["follows"][t], r["state"], q["L"], p["X"], p["Y"], q["n"], sigma=m["sigma"])
            np.testing.assert_allclose(got, us[..., mi, k], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(r["infs"][t].T, ys[..., mi, k], rtol=1e-12)
    assert lu.i == lu.v.size


@needs_ref
@pytest.mark.parametrize("tag", ["pde_coarse", "pde_default"])
def test_oracle_against_reference_pde_rhs(tag):
    if tag not in RF.manifest():
        pytest.skip("the reference's PDE part did not load when the fixtures were generated")
    p = P.pde_construct_coarse() if tag == "pde_coarse" else P.pde_construct()
    q = A.default_q()
    for sub in ("init", "pert"):
        u, z, y = (RF.load(f"{tag}_{sub}_{nm}") for nm in ("u", "z", "y"))
        du, dz, dy = P.rhs(u, z, y, p, q)
        ref = RF.load(f"{tag}_{sub}_du")
        assert np.abs(du - ref).max() <= 1e-10 * np.abs(ref).max()
        np.testing.assert_allclose(dz, RF.load(f"{tag}_{sub}_dz"), rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(dy, RF.load(f"{tag}_{sub}_dy"), rtol=1e-10, atol=1e-13)
