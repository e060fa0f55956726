"""GPU seed lifting (hb_cm_lift) against the reference's lift_plane_point outputs and the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("section", ["q3", "p3", "q2", "p2"])
@pytest.mark.parametrize("sym", [False, True])
def test_lift_bit_exact_vs_reference(section, sym):
    from hiten_b200 import centermanifold as cm
    g = np.load(os.path.join(HERE, "golden", "cm_lift.npz"))
    H = cm.PolyTable.single(g["H_deg"], g["H_coef"], g["H_exp"])
    tag = f"{section}_sym" if sym else section
    ok_ref, st_ref = g[f"ok_{tag}"].astype(bool), g[f"states_{tag}"]
    pts = g[f"pts_{section}"][: len(ok_ref)]
    ok, st = cm.lift_plane_points(H, section, pts, float(g["energy"]), symmetric=sym)
    assert np.array_equal(ok, ok_ref)
    assert np.array_equal(st[ok_ref], st_ref[ok_ref])
    assert (st[~ok_ref] == 0).all()


def test_lift_large_batch_matches_oracle_and_energy():
    """1e5 plane points: same result as the CPU oracle on a sample, and every lifted state sits on H = h0."""
    from hiten_b200 import centermanifold as cm
    g = np.load(os.path.join(HERE, "golden", "cm_lift.npz"))
    H = cm.PolyTable.single(g["H_deg"], g["H_coef"], g["H_exp"])
    Ho = O.single_poly(g["H_deg"], g["H_coef"], g["H_exp"])
    rng = np.random.default_rng(5)
    tq, tp = g["turning"][0], g["turning"][1]
    pts = np.column_stack((rng.uniform(-1.1, 1.1, 100_000) * tq, rng.uniform(-1.1, 1.1, 100_000) * tp))
    ok, st = cm.lift_plane_points(H, "p3", pts, float(g["energy"]))
    ok_o, st_o = O.cm_lift(Ho, "p3", pts[:4000], float(g["energy"]))
    assert np.array_equal(ok[:4000], ok_o.astype(bool)) and np.array_equal(st[:4000], st_o)
    assert 0.3 < ok.mean() < 0.95
    full = np.zeros((int(ok.sum()), 6))
    full[:, 1], full[:, 4], full[:, 2], full[:, 5] = st[ok, 0], st[ok, 1], st[ok, 2], st[ok, 3]
    import ctypes as C
    f = O.lib().ho_poly_eval_partial
    f.restype = C.c_double
    res = np.array([f(C.byref(Ho.struct), 0, full[i].ctypes.data_as(C.POINTER(C.c_double))) for i in range(0, len(full), 50)])
    assert np.abs(res - float(g["energy"])).max() < 1e-11
