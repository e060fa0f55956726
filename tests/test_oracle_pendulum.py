"""CPU: the reference's OWN test-suite of the Tao integrator (hiten/algorithms/integrators/_tests/test_symplectic.py),
restated for the oracle: the Taylor-pendulum Hamiltonian of its fixture, its four test configurations and its assertions
(energy conservation, reversibility on a descending grid, final-state error, RMS error / energy drift against the analytic
small-angle solution), plus bit-exactness against the trajectories the reference itself produced
(tests/golden/make_pendulum.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))

# (golden key, y0, grid, order, c_omega) -- test_symplectic.py:106-270
CASES = {
    "energy_traj": ([np.pi / 2, 0, 0, 0, 0, 0], np.linspace(0, 20.0, 2000), 6, 20.0),
    "rev_fwd": ([0.5, 0, 0, 0.3, 0, 0], np.linspace(0, 1.5, 150), 4, 5.0),
    "fse_200": ([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi, 200), 6, 5.0),
    "fse_800": ([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi, 800), 6, 5.0),
    "ivp_traj": ([0.1, 0, 0, 0, 0, 0], np.linspace(0, 100.0, 10000), 6, 20.0),
}


def hamiltonian(g, states):
    """H of the pendulum fixture evaluated from its sparse term table."""
    s = np.atleast_2d(states)
    return sum(c * np.prod(s ** e, axis=1) for c, e in zip(g["H_coef"], g["H_exp"]))


def check_reference_assertions(g, run):
    """The assertions of test_symplectic.py on trajectories produced by `run(y0, grid, order, c_omega)`."""
    tr = run(*CASES["energy_traj"])
    assert np.isclose(hamiltonian(g, tr[0]), hamiltonian(g, tr[-1]), atol=1e-5)              # :124
    fwd = run(*CASES["rev_fwd"])
    bwd = run(fwd[-1].copy(), np.linspace(1.5, 0, 150), 4, 5.0)
    assert np.allclose(CASES["rev_fwd"][0], bwd[-1], atol=1e-6)                             # :153
    assert np.allclose(run(*CASES["fse_200"])[-1], run(*CASES["fse_800"])[-1], atol=1e-5, rtol=1e-4)   # :183
    y0, grid, order, c = CASES["ivp_traj"]
    tr = run(y0, grid, order, c)
    assert np.sqrt(np.mean((tr[:, 0] - 0.1 * np.cos(grid)) ** 2)) < 0.01                    # :246-249
    e = hamiltonian(g, tr)
    assert np.max(np.abs(e - e[0])) < 1e-4                                                  # :257
    return fwd, bwd


@pytest.fixture(scope="module")
def pend():
    g = np.load(os.path.join(HERE, "golden", "pendulum.npz"))
    return g, O.PolyHam(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])


def test_fixture_table(pend):
    g, _ = pend
    # dH/dq1 = q1 - q1^3/6 + q1^5/120, dH/dp1 = p1
    assert np.diff(g["jac_ptr"]).tolist() == [3, 0, 0, 1, 0, 0]
    assert np.allclose(g["jac_coef"], [1.0, -1.0 / 6.0, 1.0 / 120.0, 1.0], rtol=1e-15)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_bit_exact_vs_reference_run(pend, name):
    g, ham = pend
    y0, grid, order, c = CASES[name]
    assert np.array_equal(O.symplectic_dense(ham, np.asarray(y0, float), grid, order, c), g[name])


def test_reference_test_suite_assertions_hold_for_the_oracle(pend):
    g, ham = pend
    fwd, bwd = check_reference_assertions(g, lambda y0, t, o, c: O.symplectic_dense(ham, np.asarray(y0, float), t, o, c))
    assert np.array_equal(fwd, g["rev_fwd"]) and np.array_equal(bwd, g["rev_bwd"])          # descending grid: dt < 0


# ---- the same fixture through the RK classes: integrators/_tests/test_rk.py:96-262 (the `_ham` kernels) -----------
RK_CASES = {
    "rk_energy_traj": ([np.pi / 6, 0, 0, 0, 0, 0], np.linspace(0, 10.0, 10000), 8),
    "rk_rev_fwd": ([0.3, 0, 0, 0.2, 0, 0], np.linspace(0, 1.0, 1000), 8),
    "rk_fse_100": ([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi / 2, 100), 6),
    "rk_fse_1600": ([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi / 2, 1600), 6),
    "rk_ivp_4": ([0.1, 0, 0, 0, 0, 0], np.linspace(0.0, 20.0, 4000), 4),
    "rk_ivp_6": ([0.1, 0, 0, 0, 0, 0], np.linspace(0.0, 20.0, 4000), 6),
    "rk_ivp_8": ([0.1, 0, 0, 0, 0, 0], np.linspace(0.0, 20.0, 4000), 8),
}


def check_reference_rk_assertions(g, run, orders=(4, 6, 8)):
    """The assertions of test_rk.py on trajectories produced by `run(y0, grid, order)`; the SciPy comparison of
    test_vs_solve_ivp is made against the analytic small-angle solution it approximates (|q| = 0.1)."""
    tr = run(*RK_CASES["rk_energy_traj"])
    e0, e1 = hamiltonian(g, tr[0])[0], hamiltonian(g, tr[-1])[0]
    assert abs(e1 - e0) / abs(e0) < 1e-6                                                     # :114-121
    fwd = run(*RK_CASES["rk_rev_fwd"])
    bwd = run(fwd[-1].copy(), np.linspace(1.0, 0, 1000), 8)
    assert np.allclose(RK_CASES["rk_rev_fwd"][0], bwd[-1], atol=1e-8, rtol=1e-6)            # :155
    assert np.allclose(run(*RK_CASES["rk_fse_100"])[-1], run(*RK_CASES["rk_fse_1600"])[-1], atol=1e-4, rtol=1e-3)  # :189
    for order in orders:
        y0, grid, _ = RK_CASES[f"rk_ivp_{order}"]
        tr = run(y0, grid, order)
        assert np.sqrt(np.mean((tr[:, 0] - g["rk_ivp_853"][:, 0]) ** 2)) < 5e-3              # :245 (DOP853 as the comparator)
        e = hamiltonian(g, tr)
        assert np.max(np.abs(e - e[0])) < 1e-2                                              # :258
    return fwd, bwd


@pytest.mark.parametrize("name", list(RK_CASES))
def test_oracle_rk_bit_exact_vs_reference_run(pend, name):
    g, ham = pend
    y0, grid, order = RK_CASES[name]
    sys_ = O.system(O.SYS_POLYHAM, ham=ham)
    assert np.array_equal(O.fixed_dense(sys_, order, np.asarray(y0, float), grid), g[name])


@pytest.mark.parametrize("method,key", [(O.RK45, "rk_ivp_45"), (O.DOP853, "rk_ivp_853")])
def test_oracle_adaptive_bit_exact_vs_reference_run(pend, method, key):
    """AdaptiveRK(order=5 | 8) with its class defaults (rtol = atol = 1e-13, max_step = inf) on the Hamiltonian system."""
    g, ham = pend
    rtol, atol, max_step, min_step = g["adaptive_defaults"]
    sys_ = O.system(O.SYS_POLYHAM, ham=ham)
    st, _ = O.adaptive_dense(sys_, method, O.HoTol(rtol, atol, max_step, min_step), np.array([0.1, 0, 0, 0, 0, 0.0]),
                             np.linspace(0.0, 20.0, 4000))
    assert np.array_equal(st, g[key])


def test_reference_rk_test_suite_assertions_hold_for_the_oracle(pend):
    g, ham = pend
    sys_ = O.system(O.SYS_POLYHAM, ham=ham)
    fwd, bwd = check_reference_rk_assertions(g, lambda y0, t, o: O.fixed_dense(sys_, o, np.asarray(y0, float), t))
    assert np.array_equal(fwd, g["rk_rev_fwd"]) and np.array_equal(bwd, g["rk_rev_bwd"])
