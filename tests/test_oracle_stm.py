"""CPU: 42-dim state+STM oracle vs the reference's _compute_stm on the halo family (BASELINE config 4)."""
import os

import numpy as np

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_var_equations_structure():
    mu = 0.0121505856
    y = np.concatenate([np.eye(6).ravel(), [0.82, 0.01, 0.03, 0.001, 0.14, -0.002]])
    d = O.var_equations(y, mu)
    F = d[:36].reshape(6, 6)           # Phi = I  ->  Phidot = F
    assert np.array_equal(F[:3, 3:], np.eye(3)) and F[3, 4] == 2.0 and F[4, 3] == -2.0
    assert np.allclose(F[3:, :3], F[3:, :3].T, rtol=0, atol=1e-15)
    assert np.allclose(d[36:], O.crtbp_accel(y[36:], mu), rtol=0, atol=1e-14)


def test_stm_family_vs_reference():
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    mu = float(g["mu"])
    s = O.system(O.SYS_VAR42, mu, fwd=1, flip=(36, 42))
    for i in range(0, 100, 9):
        y0 = np.concatenate([np.eye(6).ravel(), g["x0"][i]])
        t_eval = np.linspace(0.0, float(g["period"][i]), 2000)
        d, _ = O.adaptive_dense(s, O.DOP853, O.default_tol(), y0, t_eval)
        ref = g["PHI_end"][i]
        # not bit-exact by construction: the reference's r2**1.5 is libm pow and its 42-element np.dot is a
        # SIMD kernel; both only move the step size by ulps
        assert np.abs(d[-1][:36] - ref[:36]).max() <= 1e-10 * np.abs(ref[:36]).max()
        assert np.abs(d[-1][36:] - ref[36:]).max() <= 1e-11


def test_backward_stm_dense_rows():
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    mu = float(g["mu"])
    s = O.system(O.SYS_VAR42, mu, fwd=-1, flip=(36, 42))        # only the state block flips (rtbp.py:329)
    y0 = np.concatenate([np.eye(6).ravel(), g["x0"][0]])
    t_eval = np.linspace(0.0, float(g["period"][0]), 2000)
    d, _ = O.adaptive_dense(s, O.DOP853, O.default_tol(), y0, t_eval)
    ref = g["PHI_bwd_dense"]
    got = d[g["dense_idx"]]
    scale = np.abs(ref[:, :36]).max(axis=1, keepdims=True)
    assert (np.abs(got[:, :36] - ref[:, :36]) / scale).max() <= 1e-9
