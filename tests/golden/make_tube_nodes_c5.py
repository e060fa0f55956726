"""Manifold-tube seed data of BASELINE configs[4] for the bench workload and the full-size parity tests (from the
reference): the two tubes of examples/heteroclinic_connection.py:29-43.

For the L1 halo (Az = 0.5 southern; stable manifold, positive branch) and the L2 halo (Az = 0.3663368 northern;
unstable manifold, negative branch) this dumps, for each of the 2000 STM samples the reference snaps fractions to
(algorithms/types/services/manifold.py:470-573):
    <key>_x_node[2000,6]   state on the orbit                      (xx[idx])
    <key>_man[2000,6]      direction * Phi(t_idx) @ eigvec          (real part)
so that an initial condition is  x0W = x_node + (displacement / |man[0:3]|) * man  with tiny z, vz zeroed
(manifold.py:515-535).  The 200 fractions of the example (step = 0.005) map to node indices <key>_idx; the ICs built
from the nodes are checked against tests/golden/c5_connection.npz (the ICs the reference itself propagated).
Writes tests/golden/tube_nodes_c5.npz.   Run: python tests/golden/make_tube_nodes_c5.py  (~2 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from hiten import System  # noqa: E402


def nodes(orbit, stable, direction):
    manifold = orbit.manifold(stable=stable, direction=direction)
    svc = manifold.dynamics
    xx, tt, _, PHI = svc.compute_stm(steps=2000)
    sn, un, _ = svc.eigenvalues
    Ws, Wu, _ = svc.eigenvectors
    _, snreal_vecs = svc.stability.get_real_eigenvectors(Ws, sn)
    _, unreal_vecs = svc.stability.get_real_eigenvectors(Wu, un)
    eigvec = (snreal_vecs if svc.stable == 1 else unreal_vecs)[:, 0]
    man = np.empty((2000, 6))
    for i in range(2000):
        phi = PHI[i, :36].reshape(6, 6)
        man[i] = np.real(svc.direction * (phi @ eigvec))
    fractions = np.arange(0.0, 1.0, 0.005)
    idx = np.array([int(svc._totime(tt, f * orbit.period)[0]) for f in fractions])
    x0_ref = np.stack([svc._compute_manifold_section(period=orbit.period, fraction=f, displacement=1e-6, xx=xx, tt=tt,
                                                     PHI=PHI, eigvec=eigvec).astype(np.float64) for f in fractions])
    return np.asarray(xx), man, idx, x0_ref, np.asarray(tt), int(svc.forward)


def main():
    system = System.from_bodies("earth", "moon")
    l1 = system.get_libration_point(1)
    l2 = system.get_libration_point(2)
    halo_l1 = l1.create_orbit("halo", amplitude_z=0.5, zenith="southern")
    halo_l1.correct()
    halo_l1.propagate()
    halo_l2 = l2.create_orbit("halo", amplitude_z=0.3663368, zenith="northern")
    halo_l2.correct()
    halo_l2.propagate()
    g = np.load(os.path.join(os.path.dirname(__file__), "c5_connection.npz"))
    out = {"mu": np.float64(system.mu)}
    for key, orbit, stable, direction in (("l1", halo_l1, True, "positive"), ("l2", halo_l2, False, "negative")):
        xx, man, idx, x0_ref, tt, fwd = nodes(orbit, stable, direction)
        assert np.array_equal(x0_ref, g[f"{key}_x0W"]), f"{key}: ICs differ from the ones the reference propagated"
        assert fwd == int(g[f"{key}_forward"])
        out.update({f"{key}_x_node": xx, f"{key}_man": man, f"{key}_idx": idx, f"{key}_t_node": tt,
                    f"{key}_period": np.float64(orbit.period), f"{key}_forward": np.int64(fwd)})
        print(key, "nodes", xx.shape, "forward", fwd, "ICs of the 200 example fractions identical to c5_connection.npz")
    path = os.path.join(os.path.dirname(__file__), "tube_nodes_c5.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
