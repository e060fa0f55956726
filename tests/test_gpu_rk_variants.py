"""GPU parity: RK45 and fixed-step RK4/6/8 kernels (hb_cr3bp_rk.cu) vs the reference through the C ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "rk_variants.npz"))


def _show(name, got, ref):
    d = np.abs(np.asarray(got) - np.asarray(ref)).max()
    print(f"[parity] {name}: max |diff| {d:.2e}, bit-exact {np.array_equal(got, ref)}")
    return d


@pytest.mark.parametrize("name,fwd", [("fwd", 1), ("bwd", -1)])
def test_rk45(g, name, fwd):
    import hiten_b200 as hb
    from hiten_b200 import _lib as L
    mu, x0 = float(g["mu"]), g["x0"][None, :]
    integ = hb.make_integ(method=L.HB_RK45)
    r = hb.cr3bp_dense(x0, mu, np.linspace(0, 2.0, 41), forward=fwd, flip=(0, 6), integ=integ)
    assert _show(f"RK45 dense {name}", r.states[0], g[f"rk45_dense_{name}"]) == 0.0
    assert (r.n_acc[0], r.n_rej[0]) == (191, 1)                       # the reference's step sequence
    r = hb.cr3bp_propagate(x0, mu, 2.0, forward=fwd, flip=(0, 6), integ=integ)
    assert _show(f"RK45 final {name}", r.yf[0], g[f"rk45_final_{name}"]) == 0.0


@pytest.mark.parametrize("order", [4, 6, 8])
def test_fixed_step_bit_exact(g, order):
    import hiten_b200 as hb
    mu, x0 = float(g["mu"]), g["x0"][None, :]
    r = hb.cr3bp_dense(x0, mu, np.linspace(0, 1.0, 201), integ=hb.make_integ(method=order))
    assert np.array_equal(r.states[0][::10], g[f"rk{order}_dense"])     # no controller, IEEE ops only: bit-exact
    r = hb.cr3bp_propagate(x0, mu, 1.0, integ=hb.make_integ(method=order, n_fixed_steps=200))
    assert np.array_equal(r.yf[0], g[f"rk{order}_final"])
    r = hb.cr3bp_propagate(x0, mu, 1.0, forward=-1, flip=(0, 6), integ=hb.make_integ(method=order, n_fixed_steps=100))
    assert np.array_equal(r.yf[0], g[f"rk{order}_final_bwd"])


def test_events(g):
    import hiten_b200 as hb
    from hiten_b200 import _lib as L
    mu, T, y1 = float(g["mu"]), float(g["T"]), g["y1"][None, :]
    r = hb.cr3bp_event(y1, mu, T, 1, direction=-1, integ=hb.make_integ(method=L.HB_RK45))
    assert r.status[0] == 1 and abs(r.t_hit[0] - float(g["rk45_event_t"])) <= 1e-10
    assert _show("RK45 event state", r.yf[0], g["rk45_event_y"]) <= 1e-10
    r = hb.cr3bp_event(y1, mu, T, 0, event_offset=float(g["rk45_eventx_off"]), direction=0,
                       integ=hb.make_integ(method=L.HB_RK45))
    assert r.status[0] == 1 and abs(r.t_hit[0] - float(g["rk45_eventx_t"])) <= 1e-10
    for order in (4, 8):
        r = hb.cr3bp_event(y1, mu, T, 1, direction=-1, integ=hb.make_integ(method=order, n_fixed_steps=1500))
        assert r.status[0] == 1 and r.t_hit[0] == float(g[f"rk{order}_event_t"])
        assert np.array_equal(r.yf[0], g[f"rk{order}_event_y"])
    r = hb.cr3bp_event(y1, mu, 0.05, 1, direction=-1, integ=hb.make_integ(method=4, n_fixed_steps=10))
    assert r.status[0] == 0 and r.t_hit[0] == float(g["rk4_nohit_t"]) and np.array_equal(r.yf[0], g["rk4_nohit_y"])


def test_unsupported_method_is_an_error():
    import hiten_b200 as hb
    with pytest.raises(hb.HitenB200Error):
        hb.cr3bp_propagate(np.zeros((1, 6)) + 0.5, 0.01, 1.0, integ=hb.make_integ(method=7))
