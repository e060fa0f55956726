"""CPU: oracle tube propagation + synodic detector reproduce the reference's section hits bit for bit."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_hits(g, n_threads=4):
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    s = O.system(O.SYS_CR3BP6, mu, fwd=fwd, flip=(0, 6))
    t_eval = np.linspace(0.0, tf, steps)
    dense, _ = O.batch_dense(s, O.DOP853, O.default_tol(), g["x0W"], t_eval, n_threads)
    times = fwd * t_eval
    ht, hs, hi = [], [], []
    for i in range(len(dense)):
        t, x = O.synodic_detect(times, dense[i], 1, float(g["req_offset"]), int(g["req_direction"]), (0, 2),
                                int(g["req_segment_refine"]), float(g["req_tol_on_surface"]),
                                float(g["req_dedup_time_tol"]), float(g["req_dedup_point_tol"]))
        ht += list(t)
        hs += list(x)
        hi += [i] * len(t)
    return dense, times, np.array(hi), np.array(ht), np.array(hs)


@pytest.mark.parametrize("name,n_hits", [("c1", 121), ("c2", 679), ("se", 23)])
def test_section_hits_bit_exact(name, n_hits):
    g = np.load(os.path.join(HERE, "golden", f"synodic_{name}.npz"))
    dense, _, hi, ht, hs = oracle_hits(g)
    assert np.array_equal(dense[:, -1, :], g["yf"])
    assert len(ht) == n_hits == len(g["hit_time"])
    assert np.array_equal(hi, g["hit_traj"])
    assert np.array_equal(ht, g["hit_time"])
    assert np.array_equal(hs, g["hit_state"])
    assert np.array_equal(hs[:, [0, 2]], g["hit_point"])


def test_detector_edge_cases():
    t = np.linspace(0.0, 1.0, 11)
    x = np.zeros((11, 6))
    x[:, 1] = 0.5 - t                      # crosses y=0 at t=0.5 (a sample lies exactly on the plane)
    for direction, expect in ((-1, 1), (1, 0), (0, 1)):
        ht, hs = O.synodic_detect(t, x, 1, 0.0, direction, (0, 2), 50, 1e-6, 1e-9, 1e-6)
        assert len(ht) == expect, (direction, ht)
    ht, _ = O.synodic_detect(t[:1], x[:1], 1)            # a single sample: no segments
    assert len(ht) == 0
    ht, _ = O.synodic_detect(t, x, 1, 0.0, -1, (0, 2), 0, 1e-12, 1e-9, 1e-12)   # segment_refine = 0 path
    assert len(ht) == 1 and abs(ht[0] - 0.5) < 1e-12
    x2 = x.copy(); x2[:, 1] = 1.0 + t      # never crosses
    ht, _ = O.synodic_detect(t, x2, 1, 0.0, 0)
    assert len(ht) == 0
