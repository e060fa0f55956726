"""Oracle seed lifting against the reference's lift_plane_point outputs (tests/golden/cm_lift.npz)."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "cm_lift.npz"))


@pytest.mark.parametrize("section", ["q3", "p3", "q2", "p2"])
@pytest.mark.parametrize("sym", [False, True])
def test_lift_bit_exact(g, section, sym):
    H = O.single_poly(g["H_deg"], g["H_coef"], g["H_exp"])
    tag = f"{section}_sym" if sym else section
    ok_ref, st_ref = g[f"ok_{tag}"], g[f"states_{tag}"]
    pts = g[f"pts_{section}"][: len(ok_ref)]
    ok, st = O.cm_lift(H, section, pts, float(g["energy"]), symmetric=sym)
    assert np.array_equal(ok, ok_ref)
    assert 0 < ok.sum() < len(ok)                      # both liftable and non-liftable points are covered
    m = ok_ref.astype(bool)
    assert np.array_equal(st[m], st_ref[m])            # Brent iterates on identical residuals: bit for bit


def test_turning_points(g):
    """find_turning = solve_missing_coord with every other coordinate at zero (interfaces.py:270-295)."""
    H = O.single_poly(g["H_deg"], g["H_coef"], g["H_exp"])
    import ctypes as C
    for k, idx in enumerate((1, 4, 2, 5)):             # q2, p2, q3, p3
        root = C.c_double(0.0)
        fixed = (C.c_double * 6)(0, 0, 0, 0, 0, 0)
        ok = O.lib().ho_cm_solve_missing(C.byref(H.struct), fixed, idx, C.c_double(float(g["energy"])), C.c_double(1e-3),
                                         C.c_double(2.0), 40, 0, C.c_double(1e-12), C.byref(root))
        assert ok == 1 and root.value == g["turning"][k]
