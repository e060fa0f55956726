"""Manifold-tube seed data for the bench workload and the large-batch parity tests (from the reference).

For the BASELINE config-1 orbit (Earth-Moon L1 halo, Az=0.2 southern; stable manifold, positive branch)
this dumps, for each of the 2000 STM samples the reference snaps fractions to
(algorithms/types/services/manifold.py:470-573):
    x_node[2000,6]   state on the orbit          (xx[idx])
    man[2000,6]      direction * Phi(t_idx) @ eigvec   (real part)
so that an initial condition is  x0W = x_node + (displacement / |man[0:3]|) * man  with tiny z, vz
zeroed (manifold.py:515-535).  The 50 fractions of config 1 map to node indices `c1_idx`.
Writes tests/golden/tube_nodes_c1.npz.   Run: python tests/golden/make_tube_nodes.py  (~1.5 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from hiten import System  # noqa: E402


def main():
    system = System.from_bodies("earth", "moon")
    l1 = system.get_libration_point(1)
    halo = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
    halo.correct()
    halo.propagate()
    manifold = halo.manifold(stable=True, direction="positive")
    svc = manifold.dynamics
    xx, tt, _, PHI = svc.compute_stm(steps=2000)
    sn, un, _ = svc.eigenvalues
    Ws, Wu, _ = svc.eigenvectors
    _, snreal_vecs = svc.stability.get_real_eigenvectors(Ws, sn)
    eigvec = snreal_vecs[:, 0]
    man = np.empty((2000, 6))
    for i in range(2000):
        phi = PHI[i, :36].reshape(6, 6)
        man[i] = np.real(svc.direction * (phi @ eigvec))
    fractions = np.arange(0.0, 1.0, 0.02)
    c1_idx = np.array([int(svc._totime(tt, f * halo.period)[0]) for f in fractions])
    # cross-check with the reference's own IC routine
    x0_ref = np.stack([svc._compute_manifold_section(period=halo.period, fraction=f, displacement=1e-6, xx=xx, tt=tt,
                                                     PHI=PHI, eigvec=eigvec).astype(np.float64) for f in fractions])
    out = os.path.join(os.path.dirname(__file__), "tube_nodes_c1.npz")
    np.savez_compressed(out, mu=np.float64(system.mu), period=np.float64(halo.period), x_node=np.asarray(xx),
                        man=man, c1_idx=c1_idx, x0W_c1=x0_ref, forward=np.int64(svc.forward),
                        t_node=np.asarray(tt))
    print("wrote", out, xx.shape, man.shape)


if __name__ == "__main__":
    main()
