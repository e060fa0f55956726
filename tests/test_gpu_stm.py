"""GPU parity: the 8-lane-group 42-dim STM kernel vs the reference's _compute_stm (STMs within 1e-8)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _rel_phi(got, ref):
    return np.abs(got[..., :36] - ref[..., :36]).max(axis=-1) / np.abs(ref[..., :36]).max(axis=-1)


@pytest.mark.parametrize("arith,tol", [("parity", 1e-8), ("fast", 1e-8)])
def test_family_monodromy(arith, tol):
    import hiten_b200 as hb
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    res = hb.cr3bp_stm(g["x0"], float(g["mu"]), 0.0, tf_per_traj=g["period"], integ=hb.make_integ(arith=arith))
    assert (res.status == 0).all()
    rel = _rel_phi(res.states, g["PHI_end"])
    dx = np.abs(res.states[:, 36:] - g["PHI_end"][:, 36:]).max(axis=1)
    print(f"[parity] STM family ({arith}): |dPhi|/|Phi| median {np.median(rel):.2e} max {rel.max():.2e}; "
          f"state max {dx.max():.2e}; steps {int(res.n_acc.sum())}+{int(res.n_rej.sum())}")
    assert rel.max() <= tol           # BASELINE.json: STMs within 1e-8 (relative to |Phi|)
    assert dx.max() <= 1e-9
    assert res.n_acc[0] + res.n_rej[0] == 44          # 44 attempted steps over one halo period


def test_dense_forward_and_backward():
    import hiten_b200 as hb
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    mu, T = float(g["mu"]), float(g["period"][0])
    t_eval = np.linspace(0.0, T, 2000)
    for fwd, key in ((1, "PHI_fwd_dense"), (-1, "PHI_bwd_dense")):
        res = hb.cr3bp_stm_dense(g["x0"][:1], mu, t_eval, forward=fwd, flip=(36, 42))
        got = res.states[0][g["dense_idx"]]
        rel = _rel_phi(got, g[key])
        print(f"[parity] STM dense fwd={fwd}: max rel {rel.max():.2e}")
        assert rel.max() <= 1e-8
        assert np.array_equal(res.states[0][0, :36], np.eye(6).ravel())


def test_per_trajectory_grids_and_replicas():
    import hiten_b200 as hb
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    mu = float(g["mu"])
    x0 = np.tile(g["x0"][:8], (40, 1))                 # 320 trajectories = 10 CTAs of 32 groups
    T = np.tile(g["period"][:8], 40)
    grids = np.stack([np.linspace(0.0, t, 50) for t in T])
    a = hb.cr3bp_stm_dense(x0, mu, grids)
    b = hb.cr3bp_stm(x0, mu, 0.0, tf_per_traj=T)
    assert np.array_equal(a.states[:8], a.states[8:16])            # replicas identical
    assert np.array_equal(a.states[:, -1, :], b.states)            # dense end row == end-state call
