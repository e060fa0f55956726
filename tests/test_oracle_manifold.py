"""Oracle (oracle/ho_manifold.c) vs the reference's own outputs: manifold initial conditions and tube filters
(algorithms/types/services/manifold.py:412-424, 470-573, algorithms/common/energy.py:27-76)."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(__file__), "golden", "manifold_ics.npz")


def full_phi(g, tag):
    phi = np.zeros((g[f"{tag}_tt"].size, 42))
    phi[g[f"{tag}_rows"]] = g[f"{tag}_phi_rows"]
    return phi


@pytest.mark.parametrize("tag", ["sp", "un"])
def test_initial_conditions_bit_exact(tag):
    g = np.load(G)
    x0, idx = O.manifold_ics(full_phi(g, tag), g[f"{tag}_tt"], float(g["period"]), g[f"{tag}_eigvec"], int(g[f"{tag}_direction"]),
                             g["fractions"], g["displacements"])
    assert np.array_equal(idx, g[f"{tag}_idx"])
    assert np.array_equal(x0, g[f"{tag}_x0W"])


def test_filter_quantities_synthetic_bit_exact():
    g = np.load(G)
    out = O.tube_filter(g["syn_states"], float(g["mu"]))
    assert np.array_equal(out, g["syn_filter"], equal_nan=True)


def test_filter_quantities_on_the_default_tubes_bit_exact():
    g = np.load(G)
    mu = float(g["mu"])
    t_eval = np.linspace(0.0, float(g["c1_tf"]), int(g["c1_steps"]))
    tube, _ = O.batch_dense(O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6)), O.DOP853, O.default_tol(), g["c1_x0W"],
                            t_eval, 4)
    out = O.tube_filter(tube, mu)
    assert np.array_equal(out, g["c1_filter"])
