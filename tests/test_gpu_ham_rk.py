"""GPU parity: fixed-step RK4 / RK6 / RK8 on the polynomial Hamiltonian system (hb_ham_rk_dense / hb_ham_rk_event) vs the
reference's `_ham` kernels of `_FixedStepRK.integrate` (algorithms/integrators/rk.py:592-656, 722-757); golden vectors from
tests/golden/make_ham_rk.py, plus the oracle on a larger batch."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    from hiten_b200.centermanifold import PolyTable
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    h = np.load(os.path.join(HERE, "golden", "ham_rk.npz"))
    ham = O.PolyHam(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
    return h, PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"]), ham, g


@pytest.mark.parametrize("order", [4, 6, 8])
def test_grid_states_and_derivatives_vs_reference(gold, order):
    from hiten_b200 import symplectic as S
    h, tab, _, _ = gold
    st, der = S.integrate_rk_ham(tab, h["y0"][:4], h[f"grid_{order}"], order)
    print(f"[parity] RK{order} on the polynomial Hamiltonian: states bit-exact {np.array_equal(st, h[f'dense_{order}'])}, "
          f"derivatives bit-exact {np.array_equal(der, h[f'derivs_{order}'])}")
    assert np.array_equal(st, h[f"dense_{order}"]) and np.array_equal(der, h[f"derivs_{order}"])
    st2, none = S.integrate_rk_ham(tab, h["y0"][:4], h[f"grid_{order}"], order, want_derivatives=False)
    assert none is None and np.array_equal(st2, st)
    fast, _ = S.integrate_rk_ham(tab, h["y0"][:4], h[f"grid_{order}"], order, arith="fast")
    assert np.abs(fast - st).max() <= 1e-9


@pytest.mark.parametrize("order", [4, 6, 8])
def test_events_vs_reference(gold, order):
    from hiten_b200 import symplectic as S
    h, tab, _, _ = gold
    r = S.integrate_rk_ham_until_event(tab, h["y0"][:4], np.linspace(0.0, 6.0, 601), order, (2, 0.0, 0, 1e-12, 1e-12),
                                       want_trajectory=True)
    assert r.hit.all()
    assert np.array_equal(r.t_hit, h[f"event_{order}"][:, 0]) and np.array_equal(r.y_hit, h[f"event_{order}"][:, 1:])
    full, _ = S.integrate_rk_ham(tab, h["y0"][:4], np.linspace(0.0, 6.0, 601), order)
    for i in range(4):
        assert np.array_equal(r.traj[i, : r.n_rows[i]], full[i, : r.n_rows[i]])
    nh = S.integrate_rk_ham_until_event(tab, h["y0"][:1], np.linspace(0.0, 0.05, 6), order, (2, 10.0, 0, 1e-12, 1e-12))
    assert not nh.hit[0] and nh.t_hit[0] == h[f"nohit_{order}"][0] and np.array_equal(nh.y_hit[0], h[f"nohit_{order}"][1:])


def test_batch_vs_oracle(gold):
    from hiten_b200 import symplectic as S
    h, tab, ham, g = gold
    rng = np.random.default_rng(11)
    seeds = g["seeds_p3"][rng.integers(0, len(g["seeds_p3"]), 2000)]
    y0 = np.zeros((2000, 6))
    y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    sys_ = O.system(O.SYS_POLYHAM, ham=ham)
    t = np.sort(rng.uniform(0.0, 0.5, 33))                               # a non-uniform grid
    st, der = S.integrate_rk_ham(tab, y0, t, 8)
    for i in rng.integers(0, 2000, 24):
        assert np.array_equal(st[i], O.fixed_dense(sys_, O.RK8, y0[i], t))
        assert np.array_equal(der[i, -1], O.polyham_rhs(ham, st[i, -1]))
    ev = O.HoEvent(1, 0.0, -1, 1e-12, 1e-12)
    tl = np.linspace(0.0, 5.0, 251)
    r = S.integrate_rk_ham_until_event(tab, y0[:400], tl, 4, (1, 0.0, -1, 1e-12, 1e-12))
    for i in rng.integers(0, 400, 24):
        hit, th, yh = O.fixed_event(sys_, O.RK4, ev, y0[i], tl)
        assert hit == r.hit[i] and th == r.t_hit[i] and np.array_equal(yh, r.y_hit[i])
    with pytest.raises(ValueError):
        S.integrate_rk_ham(tab, y0[:2], t, 5)


# ---- AdaptiveRK (DOP853 / RK45) `_ham` kernels: hb_ham_adaptive_dense / _event ---------------------------------------
@pytest.mark.parametrize("order,method", [(853, 853), (45, 45)])
def test_adaptive_grid_derivatives_and_events_vs_reference(gold, order, method):
    import hiten_b200 as hb
    from hiten_b200 import symplectic as S
    h, tab, _, _ = gold
    integ = hb.make_integ(method=method, rtol=1e-11, atol=1e-12, max_step=1e300)
    r = S.integrate_adaptive_ham(tab, h["y0"][:4], h[f"grid_{order}"], integ=integ)
    assert (r.status == 0).all()
    print(f"[parity] AdaptiveRK {order} on the polynomial Hamiltonian: states bit-exact "
          f"{np.array_equal(r.states, h[f'dense_{order}'])}, derivatives bit-exact "
          f"{np.array_equal(r.derivatives, h[f'derivs_{order}'])}, steps {r.n_acc.tolist()} / {r.n_rej.tolist()}")
    assert np.array_equal(r.states, h[f"dense_{order}"]) and np.array_equal(r.derivatives, h[f"derivs_{order}"])
    e = S.integrate_adaptive_ham_until_event(tab, h["y0"][:4], 0.0, 6.0, (2, 0.0, 0, 1e-12, 1e-12), integ=integ)
    assert (e.status == 1).all()
    assert np.array_equal(e.t_hit, h[f"event_{order}"][:, 0]) and np.array_equal(e.y_hit, h[f"event_{order}"][:, 1:])
    nh = S.integrate_adaptive_ham_until_event(tab, h["y0"][:1], 0.0, 0.05, (2, 10.0, 0, 1e-12, 1e-12), integ=integ)
    assert nh.status[0] == 0 and nh.t_hit[0] == h[f"nohit_{order}"][0] and np.array_equal(nh.y_hit[0], h[f"nohit_{order}"][1:])
    fast = S.integrate_adaptive_ham(tab, h["y0"][:4], h[f"grid_{order}"],
                                    integ=hb.make_integ(method=method, arith="fast", rtol=1e-11, atol=1e-12, max_step=1e300))
    assert np.abs(fast.states - h[f"dense_{order}"]).max() <= 1e-8


def test_adaptive_batch_vs_oracle(gold):
    import hiten_b200 as hb
    from hiten_b200 import symplectic as S
    h, tab, ham, g = gold
    rng = np.random.default_rng(12)
    seeds = g["seeds_p3"][rng.integers(0, len(g["seeds_p3"]), 1500)]
    y0 = np.zeros((1500, 6))
    y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    sys_ = O.system(O.SYS_POLYHAM, ham=ham)
    t = np.linspace(0.0, 3.0, 61)
    for method, om in ((853, O.DOP853), (45, O.RK45)):
        r = S.integrate_adaptive_ham(tab, y0, t, integ=hb.make_integ(method=method), want_derivatives=False)
        assert r.derivatives is None and (r.status == 0).all()
        for i in rng.integers(0, 1500, 12):
            st, counts = O.adaptive_dense(sys_, om, O.default_tol(), y0[i], t)
            assert np.array_equal(r.states[i], st)
            assert r.n_acc[i] == counts[0] and r.n_rej[i] == counts[1]
