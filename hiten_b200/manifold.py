"""Manifold-tube host logic: initial conditions and the batched replacement of the fraction loop.

Reference: hiten/algorithms/types/services/manifold.py
  _compute_manifold_section  :470-537   (IC = orbit point + displacement along Phi*eigvec)
  _run_compute               :293-442   (serial `for fraction in ...` loop -> one batched GPU call here)
"""
import numpy as np


def manifold_initial_conditions(x_node, man, displacement):
    """x0W for every (node, displacement) pair.

    x_node[K,6]: states on the orbit; man[K,6]: direction * Phi @ eigvec (real); displacement: scalar or
    array[D].  Returns [K*D, 6] ordered displacement-major (all nodes for displacement 0, then 1, ...).
    Mirrors manifold.py:515-535: d = displacement / |man[0:3]| (1.0 if that norm < 1e-14),
    x0W = x + d * man, then z and vz are zeroed when |.| < 1e-15.
    """
    x_node = np.asarray(x_node, dtype=np.float64)
    man = np.asarray(man, dtype=np.float64)
    disp = np.atleast_1d(np.asarray(displacement, dtype=np.float64))
    mag = np.array([np.linalg.norm(m[0:3]) for m in man])
    mag = np.where(mag < 1e-14, 1.0, mag)
    out = np.empty((disp.size, x_node.shape[0], 6))
    for j, dj in enumerate(disp):
        d = dj / mag
        out[j] = x_node + d[:, None] * man
    out = out.reshape(-1, 6)
    out[np.abs(out[:, 2]) < 1.0e-15, 2] = 0.0
    out[np.abs(out[:, 5]) < 1.0e-15, 5] = 0.0
    return out
