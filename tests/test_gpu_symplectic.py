"""GPU parity: the Tao integrator over a time grid (hb_ham_symplectic_dense / _event) vs the reference's
`_ExtendedSymplectic.integrate` (algorithms/integrators/symplectic.py:877-1004); golden vectors from
tests/golden/make_symplectic.py, plus the oracle on a larger batch."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_symplectic import DENSE, EVENTS

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    from hiten_b200.centermanifold import PolyTable
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    s = np.load(os.path.join(HERE, "golden", "symplectic.npz"))
    return (s, PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"]),
            O.PolyHam(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"]), g)


@pytest.mark.parametrize("jit", [True, False], ids=["specialised", "table"])
@pytest.mark.parametrize("name", DENSE)
def test_grid_trajectories_vs_reference(gold, name, jit):
    from hiten_b200 import symplectic as S
    s, tab, _, _ = gold
    order, fwd, t0, tf, steps, c_om = s[name + "_cfg"]
    t_signed = np.linspace(t0, tf, int(steps)) * fwd
    ref = s[name]
    traj = S.integrate_symplectic(tab, s["y0"][: ref.shape[0]], t_signed, int(order), c_omega_heuristic=c_om, jit=jit)
    print(f"[parity] Tao grid {name}: {ref.shape[0]} x {int(steps)} samples, max |d| {np.abs(traj - ref).max():.2e}, "
          f"bit-exact {np.array_equal(traj, ref)}")
    assert np.array_equal(traj, ref)                                                  # bit-exact
    fast = S.integrate_symplectic(tab, s["y0"][: ref.shape[0]], t_signed, int(order), c_omega_heuristic=c_om, arith="fast",
                                  jit=jit)
    assert np.abs(fast - ref).max() <= 1e-9


@pytest.mark.parametrize("jit", [True, False], ids=["specialised", "table"])
@pytest.mark.parametrize("name", EVENTS)
def test_terminal_events_vs_reference(gold, name, jit):
    from hiten_b200 import symplectic as S
    s, tab, _, _ = gold
    order, fwd, tf, steps, idx, off, direction = s[name + "_cfg"]
    t_signed = np.linspace(0.0, tf, int(steps)) * fwd
    ref = s[name]
    r = S.integrate_symplectic_until_event(tab, s["y0"], t_signed, int(order), (int(idx), off, int(direction), 1e-12, 1e-12),
                                           want_trajectory=True, jit=jit)
    assert np.array_equal(r.hit, ref[:, 0].astype(bool))                              # identical hit flags
    assert np.array_equal(r.t_hit * fwd, ref[:, 1]) and np.array_equal(r.y_hit, ref[:, 2:])   # bit-exact
    full = S.integrate_symplectic(tab, s["y0"], t_signed, int(order))
    for i in range(len(ref)):
        assert np.array_equal(r.traj[i, : r.n_rows[i]], full[i, : r.n_rows[i]])
        assert r.n_rows[i] == (int(steps) if not r.hit[i] else r.n_rows[i]) and 1 <= r.n_rows[i] <= int(steps)
    r2 = S.integrate_symplectic_until_event(tab, s["y0"], t_signed, int(order), (int(idx), off, int(direction), 1e-12, 1e-12),
                                            jit=jit)
    assert r2.traj is None and np.array_equal(r2.t_hit, r.t_hit) and np.array_equal(r2.y_hit, r.y_hit)


def test_batch_vs_oracle_and_device_tensors(gold):
    """3000 trajectories (several CTAs per SM, the work queue refills lanes) vs the oracle, device tensors in / out."""
    import torch
    from hiten_b200 import symplectic as S
    s, tab, ham, g = gold
    rng = np.random.default_rng(5)
    seeds = g["seeds_p3"][rng.integers(0, len(g["seeds_p3"]), 3000)]
    y0 = np.zeros((3000, 6))
    y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    y0[:, 0], y0[:, 3] = 1e-3 * rng.standard_normal(3000), 1e-3 * rng.standard_normal(3000)   # off the centre manifold too
    t = np.linspace(0.0, 0.4, 41)
    out = S.integrate_symplectic(tab, torch.from_numpy(y0).cuda(), t, 4)
    assert out.is_cuda and tuple(out.shape) == (3000, 41, 6)
    out = out.cpu().numpy()
    for i in rng.integers(0, 3000, 40):
        assert np.array_equal(out[i], O.symplectic_dense(ham, y0[i], t, 4))
    ev = O.HoEvent(2, 0.0, 0, 1e-12, 1e-12)
    tl = np.linspace(0.0, 6.0, 301)
    r = S.integrate_symplectic_until_event(tab, y0[:512], tl, 4, (2, 0.0, 0, 1e-12, 1e-12))
    for i in rng.integers(0, 512, 24):
        hit, th, yh, _ = O.symplectic_event(ham, ev, y0[i], tl, 4)
        assert hit == r.hit[i] and th == r.t_hit[i] and np.array_equal(yh, r.y_hit[i])
    e = S.integrate_symplectic(tab, np.empty((0, 6)), t, 4)
    assert e.shape == (0, 41, 6)
    with pytest.raises(ValueError):
        S.integrate_symplectic(tab, y0[:4, :4], t, 4)
