"""Oracle (oracle/ho_correct.c) vs the reference's own PeriodicOrbit.correct(): Newton + Armijo
single-shooting correction of halo (analytic Jacobian + quadratic term), Lyapunov (analytic) and vertical (finite
differences) orbits; algorithms/corrector/backends/newton.py, stepping/armijo.py, operators.py:319-452."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(__file__), "golden", "correction.npz")


def test_plane_crossing_bit_exact():
    """_y_plane_crossing (singlehit/backend.py:164-282): alignment step + DOP853 event search."""
    import ctypes as C
    g = np.load(G)
    f = O.lib().ho_plane_crossing
    for x0, t_ref, x_ref in zip(g["cross_x0"], g["cross_t"], g["cross_x"]):
        t, x = C.c_double(), np.empty(6)
        hit = f(C.c_double(float(g["mu"])), O._p(np.ascontiguousarray(x0)), 1, C.c_double(0.0), C.byref(t), O._p(x))
        assert hit == 1 and t.value == t_ref and np.array_equal(x, x_ref)


@pytest.mark.parametrize("family", ["halo", "lyapunov", "vertical"])
def test_correction_matches_reference(family):
    g = np.load(G)
    xc, half, iters, rnorm, status = O.correct_orbits(g[f"{family}_x0"], float(g["mu"]), O.correct_opts(family))
    ok = g[f"{family}_iters"] >= 0
    assert np.array_equal(status == 0, ok), status           # the reference raises exactly where the oracle fails
    # The reference's convergence test (|R| < 1e-12) sits on the noise floor of its own event solver (xtol = 1e-12:
    # its halo runs stop at |R| = 9.8e-13), so a last-ulp difference in the Newton step (LAPACK vs plain elimination,
    # 42-state STM at 5e-13) can cost or save ONE iteration; everything else must agree.
    assert np.abs(iters[ok] - g[f"{family}_iters"][ok]).max() <= 1
    assert np.abs(xc[ok] - g[f"{family}_xc"][ok]).max() <= 1e-10
    assert np.abs(half[ok] - g[f"{family}_half"][ok]).max() <= 1e-10
    assert (rnorm[ok] < 1e-12).all()
    if (~ok).any():
        assert (status[~ok] == 3).all()                       # no crossing for an iterate: TypeError in the reference
    if family == "vertical":                                  # finite differences: no STM in the loop -> bit-exact
        assert np.array_equal(iters, g["vertical_iters"]) and np.array_equal(xc, g["vertical_xc"])
        assert np.array_equal(half, g["vertical_half"]) and np.array_equal(rnorm, g["vertical_rnorm"])


def test_solve_delta_reproduces_numpy_linalg_solve():
    """_solve_delta_dense's np.linalg.solve (LAPACK dgesv on OpenBLAS) for 2x2 systems, bit for bit."""
    rng = np.random.default_rng(0)
    f = O.lib().ho_solve_delta2
    for _ in range(3000):
        J = rng.standard_normal((2, 2)) * 10 ** rng.uniform(-2, 2)
        r = rng.standard_normal(2) * 10 ** rng.uniform(-6, 0)
        if np.linalg.cond(J) > 1e8:
            continue
        d = np.empty(2)
        assert f(O._p(np.ascontiguousarray(J)), O._p(r), O._p(d)) == 1
        assert np.array_equal(d, np.linalg.solve(J, -r))
