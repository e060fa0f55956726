"""Synthetic batches of BASELINE.json's configurations (bench.py, full-size parity tests, probes).

C5 (configs[4]): the two tubes of the reference's examples/heteroclinic_connection.py:29-63
  l1: L1 halo Az = 0.5 southern, stable / positive, integration_fraction 0.9, backward in time (forward = -1),
      section direction +1 (the -1 of the example flipped for the stable manifold, connections/interfaces.py:350)
  l2: L2 halo Az = 0.3663368 northern, unstable / negative, integration_fraction 1.0, forward, direction -1
  section x = 1 - mu, plane (y, z), Manifold.compute() dt = 1e-3 -> 5655 / 6284 dense samples per trajectory.
Scaled up the way SURVEY 8d prescribes: the 2000 STM nodes the reference snaps fractions to (finer `step` only
duplicates initial conditions) x displacements log-spaced in [1e-7, 1e-5], displacement-major.
The node data (tests/golden/tube_nodes_c5.npz) were dumped from the reference by tests/golden/make_tube_nodes_c5.py.
"""
import os

import numpy as np

from .manifold import manifold_initial_conditions

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GRID_DT = 1.0e-3                                   # Manifold.compute default dt (system/manifold.py:226)
C5_TUBES = {
    "l1": {"integration_fraction": 0.9, "forward": -1, "direction": 1,
           "what": "EM L1 halo Az=0.5 S, stable manifold, positive branch"},
    "l2": {"integration_fraction": 1.0, "forward": 1, "direction": -1,
           "what": "EM L2 halo Az=0.3663368 N, unstable manifold, negative branch"},
}
ENERGY_TOL = 1.0e-6                                # Manifold.compute defaults (system/manifold.py:287-288)
SAFE_DISTANCE = 2.0


def c5_nodes():
    return np.load(os.path.join(REPO, "tests", "golden", "tube_nodes_c5.npz"))


def c5_safe_radii():
    """services/manifold.py:341-345 as written (Earth-Moon distance in metres times 1e3 once more): 3.318e-05, 9.04e-06."""
    dist_m = np.float64(384400e3) * 1e3
    return SAFE_DISTANCE * (np.float64(6378.137e3) / dist_m), SAFE_DISTANCE * (np.float64(1737.4e3) / dist_m)


def c5_grid(key):
    """t_eval of one tube: linspace(0, tf, max(int(|tf| / dt) + 1, 100)) (services/manifold.py:395-397)."""
    tf = C5_TUBES[key]["integration_fraction"] * 2 * np.pi
    m = max(int(abs(tf) / GRID_DT) + 1, 100)
    return np.linspace(0.0, tf, m)


def c5_section(key, mu):
    from . import synodic
    return synodic.make_section("x", 1.0 - mu, ("y", "z"), C5_TUBES[key]["direction"])


def c5_batch(n_total, rank=0, world=1):
    """-> ({"l1": ics[n/2, 6], "l2": ics[n/2, 6]}, mu): n_total trajectories over both tubes, the rank's interleaved
    shard (i mod world).  Deterministic."""
    t = c5_nodes()
    half = n_total // 2
    out = {}
    for key in ("l1", "l2"):
        n_disp = (half + 1999) // 2000
        disp = np.logspace(-7.0, -5.0, n_disp)
        ics = manifold_initial_conditions(t[f"{key}_x_node"], t[f"{key}_man"], disp)[:half]
        out[key] = np.ascontiguousarray(ics[rank::world])
    return out, float(t["mu"])


# ---- the round-1 bench workload: BASELINE configs[0]'s tube (EM L1 halo Az = 0.2 S, stable / positive) scaled by
# displacement, with configs[1]'s section y = 0 / (x, z) / direction -1.  Kept for the large-batch parity tests.
C1_TF = 0.75 * 2.0 * np.pi


def c1_tube_batch(n, rank=0, world=1):
    """Deterministic batch: 2000 tube nodes x displacements log-spaced in [1e-7, 1e-5], interleaved over ranks."""
    t = np.load(os.path.join(REPO, "tests", "golden", "tube_nodes_c1.npz"))
    total = n * world
    n_disp = (total + 1999) // 2000
    disp = np.logspace(-7.0, -5.0, n_disp)
    ics = manifold_initial_conditions(t["x_node"], t["man"], disp)[:total]
    return np.ascontiguousarray(ics[rank::world][:n]), float(t["mu"])
