"""CPU: the C oracle against golden vectors produced by running the reference (no GPU needed)."""
import numpy as np

import oracle_lib as O


def _c1_sys(c1):
    return O.system(O.SYS_CR3BP6, float(c1["mu"]), fwd=int(c1["forward"]), flip=(0, 6))


def test_dop853_end_states_bit_exact(c1):
    """_propagate_dynsys(..., steps=2) end states (rk.py:2377-2549): bit-for-bit on all 50 trajectories."""
    yf, counts = O.batch_final(_c1_sys(c1), O.DOP853, O.default_tol(), c1["x0W"], 0.0, float(c1["tf"]), 2)
    assert np.array_equal(yf, c1["yf_steps2"])
    # trajectory 0 -> 66 accepted + 7 rejected, trajectory 25 -> 115 + 30
    assert counts[0].tolist() == [66, 7]
    assert counts[25].tolist() == [115, 30]


def test_dop853_dense_bit_exact(c1):
    """The 4713-sample dense output of Manifold.compute() at the committed sample indices."""
    t_eval = np.linspace(0.0, float(c1["tf"]), int(c1["steps"]))
    dense, _ = O.batch_dense(_c1_sys(c1), O.DOP853, O.default_tol(), c1["x0W"], t_eval, 4)
    assert np.array_equal(dense[:, c1["dense_idx"], :], c1["dense"])
    assert np.array_equal(dense[:, -1, :], c1["yf"])


def test_known_answer_events():
    """Analytic event times in the spirit of integrators/_tests/test_events.py: free fall in the
    x-direction cannot be expressed with the CR3BP field, so the KAT here is the halo's first y=0
    return: the crossing found by the terminal-event DOP853 must agree with a dense-output scan."""
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "c1_manifold.npz"))
    mu = float(g["mu"])
    s = O.system(O.SYS_CR3BP6, mu)
    y0 = g["halo_x0"].copy()
    T = float(g["halo_period"])
    ev = O.HoEvent(1, 0.0, -1, 1e-12, 1e-12)
    # start slightly after t=0 (the reference's single-hit backend does the same, singlehit/backend.py:212)
    y1, _ = O.adaptive_final(s, O.DOP853, O.default_tol(), y0, 0.0, 0.1 * T)
    hit, th, yh, _, _ = O.adaptive_event(s, O.DOP853, O.default_tol(), ev, y1, 0.0, T)
    assert hit
    assert abs(yh[1]) < 1e-11
    assert abs((th + 0.1 * T) - 0.5 * T) < 1e-6      # halo symmetry: half-period crossing
