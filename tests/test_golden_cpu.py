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
