"""CPU: polynomial-Hamiltonian oracle (RK4/6/8, Tao 2/4/6, four sections) vs the reference's _poincare_map."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("rk4_p3", 4, 0, "p3"), ("tao4_p3", 4, 1, "p3"), ("rk6_p3", 6, 0, "p3"), ("rk8_p3", 8, 0, "p3"),
         ("tao2_p3", 2, 1, "p3"), ("tao6_p3", 6, 1, "p3"), ("rk4_q3", 4, 0, "q3"), ("rk4_q2", 4, 0, "q2"),
         ("rk4_p2", 4, 0, "p2")]


@pytest.fixture(scope="module")
def cm():
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    return g, O.PolyHam(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])


@pytest.mark.parametrize("name,order,symp,sec", CASES)
def test_poincare_map_bit_exact(cm, name, order, symp, sec):
    g, ham = cm
    ref = g[name]
    seeds = g["seeds_" + sec][: len(ref)]
    f, o, t = O.cm_poincare_map(ham, seeds, float(g["dt"]), order, int(g["max_steps"]), symp, sec, float(g["c_omega"]), 4)
    assert np.array_equal(f, ref[:, 0].astype(np.int64))
    assert np.array_equal(o, ref[:, 1:5])
    assert np.array_equal(t, ref[:, 5])


def test_failed_seeds_are_flagged_not_raised(cm):
    g, ham = cm
    ref = g["rk4_p3_maxsteps200"]
    f, o, t = O.cm_poincare_map(ham, g["seeds_p3"][:64], 0.01, 4, 200, 0, "p3", 20.0, 2)
    assert np.array_equal(f, ref[:, 0].astype(np.int64)) and 0 < f.sum() < 64
    assert np.array_equal(o, ref[:, 1:5])
    assert (o[f == 0] == 0).all() and (t[f == 0] == 0).all()


def test_table_shape(cm):
    g, _ = cm
    # EM L1 degree 6 -> 29/29/34/27 non-zero gradient terms, none in d/dq1, d/dp1
    assert np.diff(g["jac_ptr"]).tolist() == [0, 29, 29, 0, 34, 27]
