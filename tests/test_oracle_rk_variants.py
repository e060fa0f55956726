"""CPU: oracle RK45 and fixed-step RK4/6/8 (dense, end state, terminal events) vs the reference, bit for bit."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "rk_variants.npz"))


@pytest.mark.parametrize("name,fwd", [("fwd", 1), ("bwd", -1)])
def test_rk45_dense_and_final(g, name, fwd):
    s = O.system(O.SYS_CR3BP6, float(g["mu"]), fwd=fwd, flip=(0, 6))
    d, _ = O.adaptive_dense(s, O.RK45, O.default_tol(), g["x0"], np.linspace(0, 2.0, 41))
    assert np.array_equal(d, g[f"rk45_dense_{name}"])
    yf, _ = O.adaptive_final(s, O.RK45, O.default_tol(), g["x0"], 0.0, 2.0)
    assert np.array_equal(yf, g[f"rk45_final_{name}"])


@pytest.mark.parametrize("order", [4, 6, 8])
def test_fixed_step(g, order):
    mu = float(g["mu"])
    d = O.fixed_dense(O.system(O.SYS_CR3BP6, mu), order, g["x0"], np.linspace(0, 1.0, 201))
    assert np.array_equal(d[::10], g[f"rk{order}_dense"]) and np.array_equal(d[-1], g[f"rk{order}_final"])
    d = O.fixed_dense(O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6)), order, g["x0"], np.linspace(0, 1.0, 101))
    assert np.array_equal(d[-1], g[f"rk{order}_final_bwd"])


def test_events(g):
    mu, T, y1 = float(g["mu"]), float(g["T"]), g["y1"]
    s = O.system(O.SYS_CR3BP6, mu)
    ev = O.HoEvent(1, 0.0, -1, 1e-12, 1e-12)
    hit, th, yh, _, _ = O.adaptive_event(s, O.RK45, O.default_tol(), ev, y1, 0.0, T)
    assert hit and th == float(g["rk45_event_t"]) and np.array_equal(yh, g["rk45_event_y"])
    ev2 = O.HoEvent(0, float(g["rk45_eventx_off"]), 0, 1e-12, 1e-12)
    hit, th, yh, _, _ = O.adaptive_event(s, O.RK45, O.default_tol(), ev2, y1, 0.0, T)
    assert hit and th == float(g["rk45_eventx_t"]) and np.array_equal(yh, g["rk45_eventx_y"])
    for order in (4, 8):
        hit, th, yh = O.fixed_event(s, order, ev, y1, np.linspace(0, T, 1501))
        assert hit and th == float(g[f"rk{order}_event_t"]) and np.array_equal(yh, g[f"rk{order}_event_y"])
    hit, th, yh = O.fixed_event(s, 4, ev, y1, np.linspace(0, 0.05, 11))
    assert not hit and th == float(g["rk4_nohit_t"]) and np.array_equal(yh, g["rk4_nohit_y"])
