"""CPU: the oracle on BASELINE configs[4] as the reference runs it (examples/heteroclinic_connection.py:29-63).

tests/golden/c5_connection.npz holds the reference's own numbers for that flow (tests/golden/make_c5.py): both tubes'
initial conditions and end states, which trajectories Manifold.compute() kept, the section hits the detection backend
returned for x = 1 - mu, (y, z) with the direction correction of connections/interfaces.py:350, and the connection
list.  The oracle must reproduce all of it bit for bit -- this geometry passes the Moon's neighbourhood (long arcs,
discarded trajectories), which configs 1-2 do not exercise."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
IDX = {"x": 0, "y": 1, "z": 2, "vx": 3, "vy": 4, "vz": 5}
ENERGY_TOL = 1e-6                           # Manifold.compute() default (system/manifold.py:287)


def c5():
    return np.load(os.path.join(HERE, "golden", "c5_connection.npz"))


def oracle_tube_hits(g, key, n_threads=8):
    """-> (dense[N,m,6], keep[N] bool, filter quantities[N,3], hits (traj among kept, t, state))."""
    mu, tf, steps, fwd = float(g["mu"]), float(g[f"{key}_tf"]), int(g[f"{key}_steps"]), int(g[f"{key}_forward"])
    s = O.system(O.SYS_CR3BP6, mu, fwd=fwd, flip=(0, 6))
    t_eval = np.linspace(0.0, tf, steps)
    dense, _ = O.batch_dense(s, O.DOP853, O.default_tol(), g[f"{key}_x0W"], t_eval, n_threads)
    q = O.tube_filter(dense, mu)
    return dense, fwd * t_eval, q


def detect_kept(g, key, dense, times, keep):
    axis = int(np.nonzero(g[f"{key}_req_normal"])[0][0])
    proj = tuple(IDX[str(c)] for c in g[f"{key}_req_plane_coords"])
    ht, hs, hi = [], [], []
    kept_idx = np.nonzero(keep)[0]
    for j, i in enumerate(kept_idx):
        t, x = O.synodic_detect(times, dense[i], axis, float(g[f"{key}_req_offset"]), int(g[f"{key}_req_direction"]),
                                proj, int(g[f"{key}_req_segment_refine"]), float(g[f"{key}_req_tol_on_surface"]),
                                float(g[f"{key}_req_dedup_time_tol"]), float(g[f"{key}_req_dedup_point_tol"]))
        ht += list(t)
        hs += list(x)
        hi += [j] * len(t)
    return np.array(hi), np.array(ht), np.array(hs).reshape(-1, 6), proj


def reference_safe_radii():
    """safe_distance (2.0) x body radius / dist_m exactly as services/manifold.py:341-345 forms them: the reference
    multiplies the Earth-Moon distance, already in metres (utils/constants.py:129), by 1e3 once more, so the radii are
    3.318e-05 and 9.04e-06 -- reproduced, not fixed."""
    dist_m = np.float64(384400e3) * 1e3
    return 2.0 * (np.float64(6378.137e3) / dist_m), 2.0 * (np.float64(1737.4e3) / dist_m)


@pytest.mark.parametrize("key", ["l1", "l2"])
def test_c5_tube_filter_and_section_hits_bit_exact(key):
    g = c5()
    dense, times, q = oracle_tube_hits(g, key)
    assert np.array_equal(dense[:, -1, :], g[f"{key}_yf"])
    r1, r2 = reference_safe_radii()
    keep = ~((q[:, 0] < r1) | (q[:, 1] < r2)) & ~(q[:, 2] > ENERGY_TOL)
    assert np.array_equal(keep, g[f"{key}_kept"])
    hi, ht, hs, proj = detect_kept(g, key, dense, times, keep)
    assert len(ht) == len(g[f"{key}_hit_time"]) > 150
    assert np.array_equal(hi, g[f"{key}_hit_traj"])
    assert np.array_equal(ht, g[f"{key}_hit_time"])
    assert np.array_equal(hs, g[f"{key}_hit_state"])
    assert np.array_equal(hs[:, list(proj)], g[f"{key}_hit_point"])


def test_c5_some_l2_trajectories_are_discarded():
    """The geometry the round-1 bench never touched: the L2 unstable tube grazes the Moon and the reference's energy
    filter drops part of it."""
    g = c5()
    assert g["l1_kept"].all() and 0 < (~g["l2_kept"]).sum() < 40


def test_c5_connections_bit_exact():
    g = c5()
    # the connection request is the two hit sets (source manifold first) in the order the synodic engine's worker
    # chunks completed (as_completed, synodic/engine.py:137: not deterministic) -- the same sets as the sorted hit lists
    for X, T, key in ((g["conn_Xu"], g["conn_tu"], "l1"), (g["conn_Xs"], g["conn_ts"], "l2")):
        a, b = np.lexsort(X.T[::-1]), np.lexsort(g[f"{key}_hit_state"].T[::-1])
        assert np.array_equal(X[a], g[f"{key}_hit_state"][b]) and np.array_equal(T[a], g[f"{key}_hit_traj"][b])
    r = O.connections(g["conn_pu"], g["conn_ps"], g["conn_Xu"], g["conn_Xs"], float(g["conn_eps"]),
                      float(g["conn_dv_tol"]), float(g["conn_bal_tol"]))
    assert r["pairs_considered"] == int(g["conn_pairs_considered"])
    assert len(r["dv"]) == len(g["conn_dv"]) == 6
    assert np.array_equal(r["iu"], g["conn_iu"]) and np.array_equal(r["is_"], g["conn_is"])
    assert np.array_equal(r["kind"], g["conn_kind"]) and np.array_equal(r["dv"], g["conn_dv"])
    assert np.array_equal(r["pt"], g["conn_pt"])
    assert np.array_equal(r["su"], g["conn_su"]) and np.array_equal(r["ss"], g["conn_ss"])
    assert np.array_equal(g["conn_tu"][r["iu"]], g["conn_tiu"]) and np.array_equal(g["conn_ts"][r["is_"]], g["conn_tis"])
