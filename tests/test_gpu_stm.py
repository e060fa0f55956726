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


def _crtbp_field(s, mu):
    x, y, z, vx, vy, vz = s.T
    r1 = np.sqrt((x + mu) ** 2 + y * y + z * z)
    r2 = np.sqrt((x - 1 + mu) ** 2 + y * y + z * z)
    ax = 2 * vy + x - (1 - mu) * (x + mu) / r1 ** 3 - mu * (x - 1 + mu) / r2 ** 3
    ay = -2 * vx + y - (1 - mu) * y / r1 ** 3 - mu * y / r2 ** 3
    az = -(1 - mu) * z / r1 ** 3 - mu * z / r2 ** 3
    return np.stack([vx, vy, vz, ax, ay, az], 1)


@pytest.mark.parametrize("arith", ["parity", "fast"])
def test_large_batch_stm_properties(arith):
    """BASELINE configs[3] scaled to the size that fills a B200 (100 family members x 1000 replicas = 1e5 trajectories
    of the 42-state system over one period each). Size-independent checks: replicas are bit-identical whatever lane
    group / queue position they ran in; Phi is symplectic in canonical coordinates (p = v + omega x r), det Phi = 1,
    and Phi(T) f(x0) = f(x(T)) (the flow maps its own vector field)."""
    import hiten_b200 as hb
    g = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
    mu, reps = float(g["mu"]), 1000
    x0 = np.tile(g["x0"], (reps, 1))
    T = np.tile(g["period"], reps)
    res = hb.cr3bp_stm(x0, mu, 0.0, tf_per_traj=T, integ=hb.make_integ(arith=arith))
    assert (res.status == 0).all()
    k = len(g["x0"])
    out = res.states.reshape(reps, k, 42)
    assert np.array_equal(out, np.broadcast_to(out[0], out.shape))
    assert np.array_equal(res.n_acc.reshape(reps, k), np.broadcast_to(res.n_acc[:k], (reps, k)))
    assert _rel_phi(out[0], g["PHI_end"]).max() <= 1e-8
    Phi, xT = out[0][:, :36].reshape(-1, 6, 6), out[0][:, 36:]
    K = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    Tm = np.block([[np.eye(3), np.zeros((3, 3))], [K, np.eye(3)]])
    J = np.block([[np.zeros((3, 3)), np.eye(3)], [-np.eye(3), np.zeros((3, 3))]])
    Pc = Tm @ Phi @ np.linalg.inv(Tm)
    scale = np.abs(Pc).max(axis=(1, 2))
    sympl = np.abs(np.swapaxes(Pc, 1, 2) @ J @ Pc - J).max(axis=(1, 2)) / scale ** 2
    det = np.abs(np.linalg.det(Phi) - 1.0)
    f0 = _crtbp_field(g["x0"], mu)
    flow = np.abs(np.einsum("nij,nj->ni", Phi, f0) - _crtbp_field(xT, mu)).max(axis=1) / (scale * np.abs(f0).max(axis=1))
    print(f"[parity] STM 1e5 batch ({arith}): symplectic residual {sympl.max():.1e}, |det - 1| {det.max():.1e}, "
          f"flow identity {flow.max():.1e}")
    assert sympl.max() <= 1e-8 and det.max() <= 1e-7 and flow.max() <= 1e-8
