"""The plain-C restatement (oracle/abm_oracle.c) against the NumPy restatement (oracle/abm_oracle.py), which is
pinned to the closed-form known answers of test_oracle_kats.py.  CPU only."""
import numpy as np

from oracle import abm_oracle as A
from oracle import c_oracle as CO
from oracle import ensemble_oracle as E


def make(seed, n, NT):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-2.2, 2.2, (n, 2))
    state = np.where(rng.random(n) > 0.5, 1, -1)
    follow0, inf0, media0 = A.abm_init(x0)
    return x0, state, follow0, inf0, media0, rng.standard_normal((NT, n, 2)), rng.random((NT, n)), rng.random((NT, n))


def test_c_step_equals_numpy_oracle():
    for n, NT, seed in ((60, 40, 1), (250, 30, 2), (7, 25, 3)):
        p, q = A.default_p(), A.default_q(n=n)
        r = make(seed, n, NT)
        a = A.abm_solve(NT, *r, p, q)
        c = CO.abm_solve(NT, *r, p, q)
        np.testing.assert_array_equal(c["follows"], a["follows"])          # discrete state: identical
        # the C loops add left to right, NumPy pairwise: agreement to rounding, not bitwise
        assert np.nanmax(np.abs(c["xs"] - a["xs"])) < 1e-12
        np.testing.assert_allclose(c["infs"], a["infs"], atol=1e-12, equal_nan=True)
        np.testing.assert_allclose(c["meds"], a["meds"], atol=1e-12)
        assert (np.diff(a["follows"], axis=0) != 0).sum() > 0 or n < 20


def test_c_kde_equals_separable_numpy_kde():
    n = 80
    p, q = A.default_p(dx=0.2), A.default_q(n=n)
    x0, state, follow0, *_ = make(5, n, 1)
    ref = E.kde_snapshot(x0, follow0, state, 4, p["X"], p["Y"], n)
    got = CO.kde(x0, follow0, state, p, q)
    assert np.abs(got - ref).max() < 1e-12 * ref.max() + 1e-15
    assert abs(got.sum() * p["dV"] - ref.sum() * p["dV"]) < 1e-12


def test_c_ensemble_bench_runs_and_conserves_mass():
    n, NT = 50, 20
    p, q = A.default_p(dx=0.2), A.default_q(n=n)
    st = E.savetimes_abm(NT, 4)
    tot, used = CO.ensemble_bench(NT, 6, p, q, st, threads=2)
    assert used >= 1
    # every snapshot integrates to ~1 (Gaussians of width 0.1 lose a little mass at the boundary)
    mass = tot * p["dV"] / (6 * 4)
    assert 0.9 < mass < 1.0001
    tot2, _ = CO.ensemble_bench(NT, 6, p, q, st, threads=1)
    assert abs(tot - tot2) < 1e-9 * abs(tot)                                # independent of the thread count
