"""CPU: oracle restatement of `_ExtendedSymplectic.integrate` (algorithms/integrators/symplectic.py:877-1004:
`_integrate_symplectic` :564-653, `_integrate_symplectic_until_event` :657-782) vs golden vectors produced by running the
reference (tests/golden/make_symplectic.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
DENSE = ["d_o4_f", "d_o4_b", "d_o2_f", "d_o6_f", "d_o8_f", "d_o4_two"]
EVENTS = ["e_q3_up", "e_q3_any", "e_q2_dn_b", "e_p2_o2", "e_nohit", "e_p3_o6"]


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    s = np.load(os.path.join(HERE, "golden", "symplectic.npz"))
    return s, O.PolyHam(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])


@pytest.mark.parametrize("name", DENSE)
def test_grid_trajectories_bit_exact(gold, name):
    s, ham = gold
    order, fwd, t0, tf, steps, c_om = s[name + "_cfg"]
    t_signed = np.linspace(t0, tf, int(steps)) * fwd                      # symplectic.py:963
    ref = s[name]
    for i in range(ref.shape[0]):
        traj = O.symplectic_dense(ham, s["y0"][i], t_signed, int(order), c_om)
        assert np.array_equal(traj, ref[i]), f"{name}[{i}]: max diff {np.abs(traj - ref[i]).max():.3e}"


@pytest.mark.parametrize("name", EVENTS)
def test_terminal_events_bit_exact(gold, name):
    s, ham = gold
    order, fwd, tf, steps, idx, off, direction = s[name + "_cfg"]
    t_signed = np.linspace(0.0, tf, int(steps)) * fwd
    ref = s[name]
    ev = O.HoEvent(int(idx), float(off), int(direction), 1e-12, 1e-12)
    for i in range(ref.shape[0]):
        hit, th, yh, traj = O.symplectic_event(ham, ev, s["y0"][i], t_signed, int(order))
        assert hit == bool(ref[i, 0])
        # the class reports t_hit * fwd on a hit, t_vals[-1] * fwd otherwise (symplectic.py:980-988)
        assert th * fwd == ref[i, 1]
        assert np.array_equal(yh, ref[i, 2:])
        if not hit:
            assert traj.shape[0] == int(steps) and np.array_equal(traj[-1], yh)


def test_event_trajectory_prefix_equals_grid_run(gold):
    s, ham = gold
    order, fwd, tf, steps, idx, off, direction = s["e_q3_up_cfg"]
    t_signed = np.linspace(0.0, tf, int(steps)) * fwd
    ev = O.HoEvent(int(idx), float(off), int(direction), 1e-12, 1e-12)
    hit, th, yh, traj = O.symplectic_event(ham, ev, s["y0"][0], t_signed, int(order))
    full = O.symplectic_dense(ham, s["y0"][0], t_signed, int(order))
    assert hit and 1 <= traj.shape[0] < int(steps)
    assert np.array_equal(traj, full[: traj.shape[0]])
    assert t_signed[traj.shape[0] - 1] <= th <= t_signed[traj.shape[0]]
