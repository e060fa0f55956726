"""GPU: the reference's own Tao-integrator test-suite (integrators/_tests/test_symplectic.py) run through
hb_ham_symplectic_dense: bit-exact against the reference's trajectories of the Taylor pendulum (a Hamiltonian in q1, p1 --
the variables the centre-manifold tables never touch), and the suite's own assertions."""
import os

import numpy as np
import pytest

from test_oracle_pendulum import CASES, check_reference_assertions

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pend():
    from hiten_b200.centermanifold import PolyTable
    g = np.load(os.path.join(HERE, "golden", "pendulum.npz"))
    return g, PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])


def _run(tab, arith="parity"):
    from hiten_b200 import symplectic as S
    return lambda y0, t, o, c: S.integrate_symplectic(tab, np.asarray(y0, float)[None, :], t, o, c_omega_heuristic=c,
                                                      arith=arith)[0]


@pytest.mark.parametrize("name", list(CASES))
def test_bit_exact_vs_reference_run(pend, name):
    g, tab = pend
    y0, grid, order, c = CASES[name]
    tr = _run(tab)(y0, grid, order, c)
    print(f"[parity] pendulum {name}: {len(grid)} samples, bit-exact {np.array_equal(tr, g[name])}")
    assert np.array_equal(tr, g[name])


@pytest.mark.parametrize("arith", ["parity", "fast"])
def test_reference_test_suite_assertions(pend, arith):
    g, tab = pend
    fwd, bwd = check_reference_assertions(g, _run(tab, arith))
    if arith == "parity":
        assert np.array_equal(fwd, g["rev_fwd"]) and np.array_equal(bwd, g["rev_bwd"])      # descending grid: dt < 0


# ---- the RK classes on the same fixture (integrators/_tests/test_rk.py) through hb_ham_rk_dense ----------------------
from test_oracle_pendulum import RK_CASES, check_reference_rk_assertions  # noqa: E402


def _run_rk(tab, arith="parity"):
    from hiten_b200 import symplectic as S
    return lambda y0, t, o: S.integrate_rk_ham(tab, np.asarray(y0, float)[None, :], t, o, arith=arith,
                                               want_derivatives=False)[0][0]


@pytest.mark.parametrize("name", list(RK_CASES))
def test_rk_bit_exact_vs_reference_run(pend, name):
    g, tab = pend
    y0, grid, order = RK_CASES[name]
    tr = _run_rk(tab)(y0, grid, order)
    print(f"[parity] pendulum {name}: {len(grid)} samples, bit-exact {np.array_equal(tr, g[name])}")
    assert np.array_equal(tr, g[name])


@pytest.mark.parametrize("arith", ["parity", "fast"])
def test_reference_rk_test_suite_assertions(pend, arith):
    g, tab = pend
    fwd, bwd = check_reference_rk_assertions(g, _run_rk(tab, arith))
    if arith == "parity":
        assert np.array_equal(fwd, g["rk_rev_fwd"]) and np.array_equal(bwd, g["rk_rev_bwd"])


@pytest.mark.parametrize("method,key", [(45, "rk_ivp_45"), (853, "rk_ivp_853")])
def test_adaptive_classes_bit_exact_vs_reference_run(pend, method, key):
    """AdaptiveRK(order=5 | 8) with the class defaults on the pendulum (test_rk.py:213-214) through hb_ham_adaptive_dense."""
    import hiten_b200 as hb
    from hiten_b200 import symplectic as S
    g, tab = pend
    rtol, atol, max_step, min_step = g["adaptive_defaults"]
    integ = hb.make_integ(method=method, rtol=rtol, atol=atol, max_step=1e300, min_step=min_step)
    r = S.integrate_adaptive_ham(tab, np.array([[0.1, 0, 0, 0, 0, 0.0]]), np.linspace(0.0, 20.0, 4000), integ=integ,
                                 want_derivatives=False)
    print(f"[parity] pendulum {key}: bit-exact {np.array_equal(r.states[0], g[key])}, steps {int(r.n_acc[0])}+{int(r.n_rej[0])}")
    assert np.array_equal(r.states[0], g[key])
